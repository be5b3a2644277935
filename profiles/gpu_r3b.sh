# round 2, session 3: where k_multitau_slicef spends its time on C2 -- without the compat phases, and one ncu --set full capture
mkdir -p gpurun_out
XPCS_SF_WARPS=24 timeout 300 python bench.py --workload c2 --no-cpu --no-e2e --steps 3 --warmup 1 --no-compat > gpurun_out/bench_c2_r3b_nocompat.json 2> gpurun_out/bench_c2_r3b_nocompat.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_c2_r3b_nocompat.json').read().strip().splitlines()[-1]); k=j['kernels']
print('nocompat ms/step %.2f'%j['ms_per_step'], j['parity']['ok'], {x:round(k[x]['ms_per_step'],2) for x in k if 'multitau' in x})
PY
XPCS_SF_WARPS=24 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_multitau_slicef' -c 1 -o gpurun_out/prof_c2_r3b -f python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_full_r3b.log 2>&1; tail -1 gpurun_out/ncu_full_r3b.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/prof_c2_r3b.ncu-rep > gpurun_out/ncu_c2_r3b.txt 2>/dev/null; cat gpurun_out/ncu_c2_r3b.txt | head -40
