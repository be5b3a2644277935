# session 7: float warp kernel with the flattened pair pass (dense path, C2)
TAG=${1:-s7d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_multitau_warp.py -m gpu -x -q 2>&1 | tail -8
summ() {
python - "$1" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    e=j.get("e2e") or {}
    print("%s value %.4g %s  ms/step %.3f  e2e %s  dominant %s frac %.4f" % (sys.argv[1], j["value"], j["unit"], j["ms_per_step"], ("%.4g (%.2f ms)"%(e["value"],e["ms_per_step"])) if e else "-", j["roofline"]["kernel"], j["roofline"]["frac"]))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1].get("ms_per_step", kv[1].get("ms_per_launch",0)))[:7]: print("   %-22s %9.3f ms  %s GB/s" % (k, v.get("ms_per_step", v.get("ms_per_launch")), v.get("gbs")))
except Exception as ex: print(sys.argv[1], "failed", ex)
PY
}
timeout 900 python bench.py --workload c2 --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err || tail -8 gpurun_out/bench_c2_$TAG.err
summ gpurun_out/bench_c2_$TAG.json
