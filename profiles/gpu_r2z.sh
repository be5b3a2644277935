mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_multitau_slice' -c 1 -o gpurun_out/prof_slice_d -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/prof_slice_d.ncu-rep > gpurun_out/ncu_slice_d.txt 2>/dev/null; cat gpurun_out/ncu_slice_d.txt
