# round 2, session 3: float slice kernel -- dense / float parity tests, then the C2 bench with either float kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_r3a.log 2>&1; tail -4 gpurun_out/pytest_r3a.log
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --workload c2 --no-cpu --no-e2e --steps 3 --warmup 1 > gpurun_out/bench_c2_r3a_$tag.json 2> gpurun_out/bench_c2_r3a_$tag.err; echo "bench $tag exit $?"; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_c2_r3a_$tag.json').read().strip().splitlines()[-1]); k=j['kernels']
    print('$tag', 'ms/step %.2f'%j['ms_per_step'], 'parity', j.get('parity'), {x:round(k[x]['ms_per_step'],2) for x in k})
except Exception as e:
    print('$tag', 'no line', e); print(open('gpurun_out/bench_c2_r3a_$tag.err').read()[-1500:])
PY
}
run slice XPCS_X=1
run warp XPCS_MTF_KERNEL=warp
run slice24 XPCS_SF_WARPS=24
