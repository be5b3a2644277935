# session 7 state: parity suite, all bench legs (c3 default with cpu baseline, c2, c4, c5), reference arm, launch list, ncu full
TAG=${1:-s7y}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
summ() {
python - "$1" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    e=j.get("e2e") or {}
    print("%s value %.4g %s  ms/step %.3f  e2e %s  dominant %s frac %.4f" % (sys.argv[1], j["value"], j["unit"], j["ms_per_step"], ("%.4g (%.2f ms)"%(e["value"],e["ms_per_step"])) if e else "-", j["roofline"]["kernel"], j["roofline"]["frac"]))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1].get("ms_per_step", kv[1].get("ms_per_launch",0))): print("   %-22s %9.3f ms  %s GB/s" % (k, v.get("ms_per_step", v.get("ms_per_launch")), v.get("gbs")))
    print("   cpu_baseline", j.get("cpu_baseline"))
except Exception as ex: print(sys.argv[1], "failed", ex)
PY
}
timeout 900 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err || tail -8 gpurun_out/bench_c3_$TAG.err
summ gpurun_out/bench_c3_$TAG.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err || tail -8 gpurun_out/bench_ref_$TAG.err
cat gpurun_out/bench_ref_$TAG.json | cut -c1-400
for W in c2 c4 c5; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err || tail -8 gpurun_out/bench_${W}_$TAG.err
  summ gpurun_out/bench_${W}_$TAG.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file gpurun_out/launches_c3_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file gpurun_out/launches_c2_$TAG.csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_multitau|k_scatter|k_finalize|k_hist|k_segment_reduce|k_concat' -c 7 -o gpurun_out/prof_c3_$TAG -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_dense_filter_vec|k_multitau_warpf|k_finalize_warp|k_scatter|k_hist|k_place' -c 6 -o gpurun_out/prof_c2_$TAG -f python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
ls -la gpurun_out | tail -12
