# round 2: whole gpu suite (log kept) + first C5 0.1 % point on one GPU
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; tail -12 gpurun_out/pytest_$TAG.log
free -g | head -2
timeout 1200 python bench.py --workload c5 --occupancy 0.001 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-rows 500 > gpurun_out/bench_c5_1e-3_$TAG.json 2> gpurun_out/bench_c5_1e-3_$TAG.err; echo "c5 exit $?"; tail -5 gpurun_out/bench_c5_1e-3_$TAG.err
python - gpurun_out/bench_c5_1e-3_$TAG.json <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    print("value %.4g ms/step %.1f events %d" % (j["value"], j["ms_per_step"], j["config"]["events_total"]))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]): print("   %-22s %9.3f ms x%.1f %s GB/s" % (k, v["ms_per_step"], v["launches_per_step"], v["gbs"]))
    print("   parity", j.get("parity"))
except Exception as ex: print("failed", ex)
PY
