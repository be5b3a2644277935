mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multitau_warp.py tests/test_gpu_parity.py tests/test_gpu_config_scale.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -4
for LD in 1 2 3 4; do
  XPCS_MW_LD=$LD timeout 300 python bench.py --no-cpu --no-e2e --steps 5 > gpurun_out/bench_c3_ld$LD.json 2>/dev/null
  python - $LD <<'PY'
import json,sys
j=json.load(open("gpurun_out/bench_c3_ld%s.json"%sys.argv[1])); print("ld_factor", sys.argv[1], "ms/step %.3f multitau %.3f parity %s" % (j["ms_per_step"], j["kernels"]["k_multitau_warp"]["ms_per_step"], j["parity"]["ok"]))
PY
done
XPCS_MW_LD=2 timeout 300 python bench.py --workload c1 --no-cpu --no-e2e --steps 5 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('c1 ld2', j['ms_per_step'], j['kernels']['k_multitau_warp']['ms_per_step'])"
XPCS_MW_LD=4 timeout 300 python bench.py --workload c1 --no-cpu --no-e2e --steps 5 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('c1 ld4', j['ms_per_step'], j['kernels']['k_multitau_warp']['ms_per_step'])"
