mkdir -p gpurun_out
bash profiles/gpu_sweep.sh 1 r2h "0.001"
# one GPU's share of the 1 % point at 8 GPUs: 724x724 pixels (~410 k rows) x 1 M frames x 1 % = 5.2e9 events, rows of 10^4 events
timeout 900 python bench.py --workload c5 --hw 724 724 --occupancy 0.01 --steps 1 --warmup 1 --no-e2e --no-cpu --parity-rows 32 > gpurun_out/c5_share_1pct_r2h.json 2> gpurun_out/c5_share_1pct_r2h.err; echo "exit $?"; tail -5 gpurun_out/c5_share_1pct_r2h.err | cut -c1-400
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/c5_share_1pct_r2h.json"))
    print("value %.4g ms/step %.1f events %.3g parity %s" % (j["value"], j["ms_per_step"], j["config"]["events_total"], j.get("parity")))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:9]: print("     %-20s %9.3f ms" % (k, v["ms_per_step"]))
except Exception as ex: print("failed", ex)
PY
nvidia-smi --query-gpu=memory.used --format=csv,noheader
