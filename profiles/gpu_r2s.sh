mkdir -p gpurun_out
timeout 600 python profiles/corr_wall.py c1 > gpurun_out/corr_wall_c1.json 2> gpurun_out/corr_wall_c1.err; tail -2 gpurun_out/corr_wall_c1.err
timeout 900 python profiles/corr_wall.py c3 > gpurun_out/corr_wall_c3.json 2> gpurun_out/corr_wall_c3.err; tail -2 gpurun_out/corr_wall_c3.err
python - <<'PY'
import json
for k in ("c1","c3"):
    try:
        j=json.load(open("gpurun_out/corr_wall_%s.json"%k)); print(k, "best wall %.3f s -> %.0f frames/s"%(j["best_wall_s"], j["frames_per_s"])); print("   ", j["runs"][-1]["scopes_s"])
    except Exception as e: print(k, "failed", e)
PY
