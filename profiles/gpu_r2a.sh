# round 2, first GPU call: the whole gpu suite (incl. config-scale C1 parity, StaticMap two-time, Rigaku stride/avg,
# slab partition on one GPU), smoke, and the default bench line with its parity block
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
summ() {
python - "$1" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    e=j.get("e2e") or {}
    print("%s value %.4g %s  ms/step %.3f  e2e %s  dominant %s frac %.4f" % (sys.argv[1], j["value"], j["unit"], j["ms_per_step"], ("%.4g (%.2f ms)"%(e["value"],e["ms_per_step"])) if e else "-", j["roofline"]["kernel"], j["roofline"]["frac"]))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1].get("ms_per_step", kv[1].get("ms_per_launch",0))): print("   %-22s %9.3f ms  %s GB/s" % (k, v.get("ms_per_step", v.get("ms_per_launch")), v.get("gbs")))
    print("   parity", j.get("parity"))
    print("   cpu_baseline", j.get("cpu_baseline"))
except Exception as ex: print(sys.argv[1], "failed", ex)
PY
}
timeout 900 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "bench exit $?"; tail -5 gpurun_out/bench_c3_$TAG.err
summ gpurun_out/bench_c3_$TAG.json
timeout 600 python bench.py --workload c1 --steps 5 > gpurun_out/bench_c1_$TAG.json 2> gpurun_out/bench_c1_$TAG.err; echo "bench c1 exit $?"; tail -3 gpurun_out/bench_c1_$TAG.err
summ gpurun_out/bench_c1_$TAG.json
