# round 2, session 3: k_multitau_slicef, pair pieces first in the queue -- first dense level / dense pieces / pair pieces on C2
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --workload c2 --no-cpu --no-e2e --steps 3 --warmup 1 > gpurun_out/bench_c2_r3d_$tag.json 2> gpurun_out/bench_c2_r3d_$tag.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_c2_r3d_$tag.json').read().strip().splitlines()[-1]); k=j['kernels']
    print('$tag', 'ms/step %.2f'%j['ms_per_step'], 'parity', j['parity']['ok'], '%.2g'%j['parity']['worst_rel_err'], {x:round(k[x]['ms_per_step'],2) for x in k if 'multitau' in x})
except Exception as e:
    print('$tag', 'no line', e); print(open('gpurun_out/bench_c2_r3d_$tag.err').read()[-1500:])
PY
}
run ld4_nd24 XPCS_X=1
run ld4_nd6 XPCS_SF_DENSE_PIECES=6
run ld8_nd12 XPCS_SF_LD=8 XPCS_SF_DENSE_PIECES=12
run ld8_nd24 XPCS_SF_LD=8
run ld8_nd24_p12 XPCS_SF_LD=8 XPCS_SF_PAIR_PIECES=9 XPCS_SF_PAIR_TAIL=3
run ld16_nd24_p12 XPCS_SF_LD=16 XPCS_SF_PAIR_PIECES=9 XPCS_SF_PAIR_TAIL=3
