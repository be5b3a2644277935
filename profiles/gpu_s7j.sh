# session 7: pipelined sparse ingest; pipeline tests first, then the whole suite and the c3/c5 benches (with e2e)
TAG=${1:-s7j}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
summ() {
python - "$1" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    e=j.get("e2e") or {}
    print("%s value %.4g %s  ms/step %.3f  e2e %s  dominant %s frac %.4f" % (sys.argv[1], j["value"], j["unit"], j["ms_per_step"], ("%.4g (%.2f ms)"%(e["value"],e["ms_per_step"])) if e else "-", j["roofline"]["kernel"], j["roofline"]["frac"]))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1].get("ms_per_step", kv[1].get("ms_per_launch",0)))[:6]: print("   %-22s %9.3f ms  %s GB/s" % (k, v.get("ms_per_step", v.get("ms_per_launch")), v.get("gbs")))
except Exception as ex: print(sys.argv[1], "failed", ex)
PY
}
for W in c3 c5; do
  timeout 900 python bench.py --workload $W --no-cpu --steps 5 --warmup 3 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err || tail -8 gpurun_out/bench_${W}_$TAG.err
  summ gpurun_out/bench_${W}_$TAG.json
done
XPCS_NO_PIPELINE=1 timeout 900 python bench.py --workload c3 --no-cpu --steps 5 --warmup 3 > gpurun_out/bench_c3_${TAG}_nopipe.json 2> gpurun_out/bench_c3_${TAG}_nopipe.err
summ gpurun_out/bench_c3_${TAG}_nopipe.json
