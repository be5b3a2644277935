# round 2, session 3: whole GPU suite, smoke, default bench line, C2 with the IF / IP parts before / after the dense pieces
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r3g.log 2>&1; tail -3 gpurun_out/pytest_r3g.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --workload c2 --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/bench_c2_r3g_$tag.json 2> gpurun_out/bench_c2_r3g_$tag.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_c2_r3g_$tag.json').read().strip().splitlines()[-1]); k=j['kernels']
    print('$tag', 'ms/step %.2f'%j['ms_per_step'], 'parity', j['parity']['ok'], '%.2g'%j['parity']['worst_rel_err'], {x:round(k[x]['ms_per_step'],2) for x in k if 'multitau' in x})
except Exception as e:
    print('$tag', 'no line', e); print(open('gpurun_out/bench_c2_r3g_$tag.err').read()[-1500:])
PY
}
run iofirst XPCS_X=1
run iolast XPCS_SF_IO_FIRST=0
timeout 600 python bench.py > gpurun_out/bench_c3_r03.json 2> gpurun_out/bench_c3_r03.err; echo "bench c3 exit $?"
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_c3_r03.json').read().strip().splitlines()[-1])
print('c3 ms/step %.3f value %.4g e2e %.2f ms parity %s roof %.4f cpu %s'%(j['ms_per_step'], j['value'], j['e2e']['ms_per_step'], j['parity']['ok'], j['roofline']['frac'], j.get('cpu_baseline',{}).get('value')))
PY
