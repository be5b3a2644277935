mkdir -p gpurun_out
python bench.py --workload c4 --steps 3 --warmup 1 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err || tail -5 gpurun_out/bench_c4.err
cut -c1-1800 gpurun_out/bench_c4.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_twotime_gemm -s 1 -c 1 -o gpurun_out/prof_c4 -f python bench.py --workload c4 --steps 1 --warmup 1 > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
