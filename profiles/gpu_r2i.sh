TAG=${1:-r2i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; tail -8 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "c3 exit $?"
timeout 900 python bench.py --workload c2 --steps 3 --warmup 2 --no-e2e > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "c2 exit $?"; tail -2 gpurun_out/bench_c2_$TAG.err
python - $TAG <<'PY'
import json,sys
for w in ("c3","c2"):
    f="gpurun_out/bench_%s_%s.json"%(w,sys.argv[1])
    try:
        j=json.load(open(f)); print(f, "value %.4g ms/step %.3f" % (j["value"], j["ms_per_step"]), "e2e", (j.get("e2e") or {}).get("ms_per_step"), "parity", (j.get("parity") or {}).get("ok"), "cpu", (j.get("cpu_baseline") or {}).get("value"))
        for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:7]: print("     %-20s %9.3f ms" % (k, v["ms_per_step"]))
    except Exception as ex: print(f, "failed", ex)
PY
