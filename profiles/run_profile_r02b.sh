# round 2, final state (k_multitau_slice): default bench line with cpu_baseline, C1 / C5 lines, launch list
# (gpu__time_duration) of the default bench command, one ncu --set full capture of the kernels of the C3 step
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "bench exit $?"
timeout 900 python bench.py --workload c1 --steps 5 > gpurun_out/bench_c1_$TAG.json 2> gpurun_out/bench_c1_$TAG.err; echo "c1 exit $?"
timeout 900 python bench.py --workload c5 --steps 3 --no-cpu > gpurun_out/bench_c5_$TAG.json 2> gpurun_out/bench_c5_$TAG.err; echo "c5 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-parity > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_multitau_slice|k_scatter_rec|k_finalize|k_hist|k_segment_reduce' -c 5 -o gpurun_out/prof_c3_$TAG -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/prof_c3_$TAG.ncu-rep > gpurun_out/ncu_c3_$TAG.txt 2>/dev/null; grep -c Kernel gpurun_out/ncu_c3_$TAG.txt
ls -la gpurun_out/*$TAG* | head -20
