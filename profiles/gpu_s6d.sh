# session 6: dense filter v3 (SIMD max prefilter, private slots, per-warp flush)
TAG=${1:-s6d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
summ() {
python - "$1" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    e=j.get("e2e") or {}
    print("%s value %.4g %s  ms/step %.3f  e2e %s  dominant %s frac %.4f" % (sys.argv[1], j["value"], j["unit"], j["ms_per_step"], ("%.4g (%.2f ms)"%(e["value"],e["ms_per_step"])) if e else "-", j["roofline"]["kernel"], j["roofline"]["frac"]))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1].get("ms_per_step", kv[1].get("ms_per_launch",0))): print("   %-22s %9.3f ms  %s GB/s" % (k, v.get("ms_per_step", v.get("ms_per_launch")), v.get("gbs")))
except Exception as ex: print(sys.argv[1], "failed", ex)
PY
}
timeout 900 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err || tail -8 gpurun_out/bench_c2_$TAG.err
summ gpurun_out/bench_c2_$TAG.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_dense_filter_vec|k_multitau_warpf|k_finalize_warp' -c 3 -o gpurun_out/prof_c2_$TAG -f python bench.py --workload c2 --frames 8000 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log
