# session 7: stream scatter + placement; all parity suites, benches c3/c5/c2
TAG=${1:-s7f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
summ() {
python - "$1" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    e=j.get("e2e") or {}
    print("%s value %.4g %s  ms/step %.3f  e2e %s  dominant %s frac %.4f" % (sys.argv[1], j["value"], j["unit"], j["ms_per_step"], ("%.4g (%.2f ms)"%(e["value"],e["ms_per_step"])) if e else "-", j["roofline"]["kernel"], j["roofline"]["frac"]))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1].get("ms_per_step", kv[1].get("ms_per_launch",0)))[:8]: print("   %-22s %9.3f ms  %s GB/s" % (k, v.get("ms_per_step", v.get("ms_per_launch")), v.get("gbs")))
except Exception as ex: print(sys.argv[1], "failed", ex)
PY
}
for W in c3 c5 c2; do
  X=""; [ $W = c2 ] && X="--no-e2e"
  timeout 900 python bench.py --workload $W --no-cpu $X --steps 3 --warmup 3 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err || tail -8 gpurun_out/bench_${W}_$TAG.err
  summ gpurun_out/bench_${W}_$TAG.json
done
