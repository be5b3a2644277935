# round 2, session 3: k_multitau_slicef compiled for up to 32 warps (64 registers, 24 bytes of spill) -- warps / pieces on C2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -x -q > gpurun_out/pytest_r3f.log 2>&1; tail -2 gpurun_out/pytest_r3f.log
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --workload c2 --no-cpu --no-e2e --steps 3 --warmup 1 > gpurun_out/bench_c2_r3f_$tag.json 2> gpurun_out/bench_c2_r3f_$tag.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_c2_r3f_$tag.json').read().strip().splitlines()[-1]); k=j['kernels']
    print('$tag', 'ms/step %.2f'%j['ms_per_step'], 'parity', j['parity']['ok'], '%.2g'%j['parity']['worst_rel_err'], {x:round(k[x]['ms_per_step'],2) for x in k if 'multitau' in x})
except Exception as e:
    print('$tag', 'no line', e); print(open('gpurun_out/bench_c2_r3f_$tag.err').read()[-1500:])
PY
}
run w24 XPCS_X=1
run w32 XPCS_SF_WARPS=32
run w32_nd16 XPCS_SF_WARPS=32 XPCS_SF_DENSE_PIECES=16
run w32_nd16_p12 XPCS_SF_WARPS=32 XPCS_SF_DENSE_PIECES=16 XPCS_SF_PAIR_PIECES=9 XPCS_SF_PAIR_TAIL=3
