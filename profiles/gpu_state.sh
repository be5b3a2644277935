# full state check: parity suite, the three bench legs, launch list, ncu full of the top kernels
# usage: bash profiles/gpu_state.sh <tag>
TAG=${1:-state}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for W in c3 c2 c4; do
  timeout 900 python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err || tail -8 gpurun_out/bench_${W}_$TAG.err
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
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file gpurun_out/launches_c3_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_multitau|k_scatter|k_finalize|k_hist|k_segment_reduce' -c 6 -o gpurun_out/prof_c3_$TAG -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
