mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_multitau_slice' -c 1 -o gpurun_out/prof_slice_b -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/prof_slice_b.ncu-rep > gpurun_out/ncu_slice_b.txt 2>/dev/null; cat gpurun_out/ncu_slice_b.txt
timeout 300 python bench.py --no-cpu --no-e2e --steps 5 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); k=j['kernels']; print('c3 ms/step %.3f parity %s'%(j['ms_per_step'], j['parity']['ok']), {x:round(k[x]['ms_per_step'],3) for x in k if 'multitau' in x})"
