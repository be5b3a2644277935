# multi-GPU check of the final state: N = $1 GPUs: multi-GPU tests (N >= 2) and the strong-scaling bench line
n=${1:-2}; TAG=r02b
mkdir -p gpurun_out
if [ "$n" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -q > gpurun_out/pytest_multigpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_multigpu_$TAG.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c3_${n}gpu_$TAG.json 2> gpurun_out/bench_c3_${n}gpu_$TAG.err; echo "bench N=$n exit $?"; tail -2 gpurun_out/bench_c3_${n}gpu_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c3_${n}gpu_$TAG.json'))
print('N=%d ms/step %.3f value %.4g e2e %.2f ms parity %s'%(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['parity']))
PY
