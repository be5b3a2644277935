set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err
python bench.py --workload c1 --steps 5 --warmup 3 --cpu-frames 2000 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -3 gpurun_out/bench_c1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_multitau|k_scatter|k_finalize|k_hist|k_segment_reduce' -s 5 -c 5 -o gpurun_out/prof_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
