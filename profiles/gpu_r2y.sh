mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multitau_warp.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_r2y.log 2>&1; tail -2 gpurun_out/pytest_r2y.log
run() { env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 5 $EXTRA 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); k=j['kernels']; print('$* $EXTRA: c3 ms/step %.3f parity %s'%(j['ms_per_step'], j['parity']['ok']), {x:round(k[x]['ms_per_step'],3) for x in k if 'multitau' in x})"; }
run XPCS_SL_PAIR_TAIL=8
run XPCS_SL_PAIR_TAIL=8 XPCS_SL_IO_PIECES=3
run XPCS_SL_PAIR_TAIL=8 XPCS_SL_IO_PIECES=4
EXTRA=--no-compat run XPCS_SL_PAIR_TAIL=8
