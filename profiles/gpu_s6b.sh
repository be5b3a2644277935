# session 6: vectorised dense filter + float warp-per-row multi-tau: parity suite, c2 legs, finalize smem experiment
TAG=${1:-s6b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dense.py 2>&1 | tail -5
summ() {
python - "$1" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    e=j.get("e2e") or {}
    print("%s value %.4g %s  ms/step %.3f  e2e %s  dominant %s frac %.4f" % (sys.argv[1], j["value"], j["unit"], j["ms_per_step"], ("%.4g (%.2f ms)"%(e["value"],e["ms_per_step"])) if e else "-", j["roofline"]["kernel"], j["roofline"]["frac"]))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1].get("ms_per_step", kv[1].get("ms_per_launch",0))): print("   %-22s %9.3f ms" % (k, v.get("ms_per_step", v.get("ms_per_launch"))))
except Exception as ex: print(sys.argv[1], "failed", ex)
PY
}
timeout 900 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err || tail -8 gpurun_out/bench_c2_$TAG.err
summ gpurun_out/bench_c2_$TAG.json
for KB in 24 64; do
XPCS_FIN_SMEM_KB=$KB timeout 900 python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c2_fin$KB.json 2> gpurun_out/bench_c2_fin$KB.err || tail -8 gpurun_out/bench_c2_fin$KB.err
summ gpurun_out/bench_c2_fin$KB.json
done
timeout 900 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err || tail -8 gpurun_out/bench_c3_$TAG.err
summ gpurun_out/bench_c3_$TAG.json
