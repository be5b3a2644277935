# round 2, session 3: k_multitau_slicef with the cheaper dense walk -- dense tests, then the number of dense pieces on C2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -x -q > gpurun_out/pytest_r3c.log 2>&1; tail -3 gpurun_out/pytest_r3c.log
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --workload c2 --no-cpu --no-e2e --steps 3 --warmup 1 > gpurun_out/bench_c2_r3c_$tag.json 2> gpurun_out/bench_c2_r3c_$tag.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_c2_r3c_$tag.json').read().strip().splitlines()[-1]); k=j['kernels']
    print('$tag', 'ms/step %.2f'%j['ms_per_step'], 'parity', j['parity']['ok'], j['parity']['worst_rel_err'], {x:round(k[x]['ms_per_step'],2) for x in k if 'multitau' in x})
except Exception as e:
    print('$tag', 'no line', e); print(open('gpurun_out/bench_c2_r3c_$tag.err').read()[-1500:])
PY
}
run nd24 XPCS_X=1
run nd48 XPCS_SF_DENSE_PIECES=48
run nd12 XPCS_SF_DENSE_PIECES=12
