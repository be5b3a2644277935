TAG=r2g
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$TAG.log 2>&1; tail -5 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "c3 exit $?"
XPCS_FIN_WARP=1 XPCS_PLACE_KERNEL=1 timeout 600 python bench.py --no-cpu --no-e2e --steps 5 > gpurun_out/bench_c3_finwarp_$TAG.json 2> gpurun_out/bench_c3_finwarp_$TAG.err; echo "c3 finwarp exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_c3_r2g.json", "gpurun_out/bench_c3_finwarp_r2g.json"):
    try:
        j=json.load(open(f)); print(f, "ms/step %.3f" % j["ms_per_step"], "e2e", (j.get("e2e") or {}).get("ms_per_step"), "parity", (j.get("parity") or {}).get("ok"))
        for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:7]: print("     %-20s %9.3f ms" % (k, v["ms_per_step"]))
    except Exception as ex: print(f, "failed", ex)
PY
bash profiles/gpu_sweep.sh 1 $TAG "0.0001"
