# round 2, session 3, final state of the float path: ncu launch list of the C2 step and one ncu --set full capture of k_multitau_slicef
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_r03.csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_launches_r03.log 2>&1; tail -1 gpurun_out/ncu_launches_r03.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_multitau_slicef' -c 1 -o gpurun_out/prof_c2_r03 -f python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_full_r03.log 2>&1; tail -1 gpurun_out/ncu_full_r03.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/prof_c2_r03.ncu-rep > gpurun_out/ncu_c2_r03.txt 2>/dev/null; head -30 gpurun_out/ncu_c2_r03.txt
