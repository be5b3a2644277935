# round 2: multi-GPU call (gpurun --gpus N): threads-per-GPU tests, corr --gpus, and the strong-scaling bench under torchrun
N=${1:-2}
TAG=${2:-r2c}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_parity.py::test_shards_sum_to_single_gpu_partials tests/test_gpu_pipeline.py::test_pipelined_shards_sum_to_single_gpu_partials -m gpu -q 2>&1 | tail -30
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_c3_${n}gpu_$TAG.json 2> gpurun_out/bench_c3_${n}gpu_$TAG.err; echo "bench N=$n exit $?"; tail -4 gpurun_out/bench_c3_${n}gpu_$TAG.err
    python - gpurun_out/bench_c3_${n}gpu_$TAG.json <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1])); e=j["e2e"]
    print("N=%d value %.4g ms/step %.3f (host wall %.3f) e2e %.4g (%.2f ms) h2d/gpu %.1f MB scaling %s" % (j["n_gpus"], j["value"], j["ms_per_step"], j["pipeline"]["host_wall_ms_per_step"], e["value"], e["ms_per_step"], e["h2d_bytes_per_step_per_gpu"]/1e6, j["scaling"]))
    for k,v in sorted(j["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]): print("   %-22s %9.3f ms x%.1f" % (k, v["ms_per_step"], v["launches_per_step"]))
    print("   parity", j.get("parity"))
except Exception as ex: print(sys.argv[1], "failed", ex)
PY
  fi
done
