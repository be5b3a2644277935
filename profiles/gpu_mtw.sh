# new multi-tau kernel: its tests first, then the whole parity suite, then the c3 bench leg
TAG=${1:-mtw}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multitau_warp.py -x -q 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err || tail -8 gpurun_out/bench_c3_$TAG.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_c3_$TAG.json"))
e=j.get("e2e") or {}
print("c3 value %.4g ms/step %.3f e2e %.4g (%.2f ms) dominant %s frac %s" % (j["value"], j["ms_per_step"], e["value"], e["ms_per_step"], j["roofline"]["kernel"], j["roofline"]["frac"]))
for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]): print("   %-22s %9.3f ms" % (k, v["ms_per_step"]))
PY
