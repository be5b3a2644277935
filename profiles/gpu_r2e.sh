# round 2: two-time (wide tiles + rasterised tile list, StaticMap), C4 bench + ncu of the GEMM, full-size reference runs
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_twotime.py tests/test_gpu_corr_host.py -m gpu -q 2>&1 | tail -15
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err; echo "c4 exit $?"; tail -3 gpurun_out/bench_c4_$TAG.err
python - gpurun_out/bench_c4_$TAG.json <<'PY'
import json,sys
j=json.load(open(sys.argv[1])); r=j["roofline"]
print("c4 ms/step %.3f gemm %.3f ms algorithmic %.1f TF (frac %.3f, of sustained %s) issued %.1f TF" % (j["ms_per_step"], r["ms_per_launch"], r["achieved"], r["frac"], r.get("frac_of_sustained"), r["achieved_issued"]))
print(j["kernels"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_twotime_gemm -c 1 -o gpurun_out/prof_c4_$TAG -f python bench.py --workload c4 --steps 1 --warmup 0 > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log
python profiles/ncu_summary.py gpurun_out/prof_c4_$TAG.ncu-rep 2>/dev/null | head -30; ncu -i gpurun_out/prof_c4_$TAG.ncu-rep --page raw --csv 2>/dev/null | python -c "import csv,sys; r=list(csv.reader(sys.stdin)); h=r[0]; [print(k, r[2][h.index(k)]) for k in h if \"tensor\" in k or \"pipe_tensor\" in k][:12]"
timeout 1500 python profiles/ref_full_config.py c1 c3 > gpurun_out/ref_full_config_$TAG.json 2> gpurun_out/ref_full_config_$TAG.err; tail -3 gpurun_out/ref_full_config_$TAG.err
