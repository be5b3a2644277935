# usage: bash profiles/gpu_check.sh [tag]   -- parity tests, a short bench, launch list, one ncu full capture
TAG=${1:-check}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err || tail -5 gpurun_out/bench_c3_$TAG.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_c3_$TAG.json"))
print("value %.4g frames/s  ms/step %.3f  e2e %.4g (%.3f ms)  launches %d" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["gpu_launches"]))
for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]): print("  %-22s %8.3f ms  %s GB/s" % (k, v["ms_per_step"], ("%.0f"%v["gbs"]) if v["gbs"] else "-"))
print(j["roofline"]); print(j.get("cpu_baseline"))
PY
if [ "$2" = "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches_c3_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${3:-k_multitau}" -s ${4:-1} -c ${5:-1} -o gpurun_out/prof_c3_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi
