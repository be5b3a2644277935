TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$TAG.log 2>&1; tail -4 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "c3 exit $?"
python - $TAG <<'PY'
import json,sys
j=json.load(open("gpurun_out/bench_c3_%s.json"%sys.argv[1])); print("ms/step %.3f e2e %.3f parity %s" % (j["ms_per_step"], j["e2e"]["ms_per_step"], j["parity"]["ok"]))
for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:6]: print("     %-20s %9.3f ms" % (k, v["ms_per_step"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|nccl' -c 300 --csv --log-file gpurun_out/launches_c3_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-parity > gpurun_out/ncu_launches.log 2>&1; grep -c k_ gpurun_out/launches_c3_$TAG.csv
