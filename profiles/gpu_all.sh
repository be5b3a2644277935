# parity suite + the three bench legs; usage: bash profiles/gpu_all.sh <tag>
TAG=${1:-all}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for W in c3 c2 c4; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err || tail -8 gpurun_out/bench_${W}_$TAG.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/bench_${W}_$TAG.json"))
    e=j.get("e2e") or {}
    print("$W value %.4g %s  ms/step %.3f  e2e %s  dominant %s frac %s" % (j["value"], j["unit"], j["ms_per_step"], ("%.4g (%.2f ms)"%(e["value"],e["ms_per_step"])) if e else "-", j["roofline"]["kernel"], j["roofline"]["frac"]))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1].get("ms_per_step", kv[1].get("ms_per_launch",0))): print("   %-22s %9.3f ms" % (k, v.get("ms_per_step", v.get("ms_per_launch"))))
    print("   cpu_baseline", j.get("cpu_baseline"))
except Exception as ex: print("$W failed", ex)
PY
done
