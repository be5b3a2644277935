mkdir -p gpurun_out
timeout 1200 python bench.py --workload c2 --steps 3 --warmup 1 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err || tail -15 gpurun_out/bench_c2.err
cut -c1-3500 gpurun_out/bench_c2.json
