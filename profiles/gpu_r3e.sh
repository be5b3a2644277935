# round 2, session 3: cycle trace of k_multitau_slicef on C2 (library built with -DXPCS_SL_TRACE, swapped in on the box only)
mkdir -p gpurun_out
cp xpcs-eigen_b200/libxpcs_b200.so /tmp/product.so
cp xpcs-eigen_b200/libxpcs_b200_trace.so xpcs-eigen_b200/libxpcs_b200.so
timeout 300 python bench.py --workload c2 --no-cpu --no-e2e --no-parity --steps 1 --warmup 1 2>/dev/null | grep ^SLT > gpurun_out/trace_slicef_ld4.txt
XPCS_SF_LD=8 XPCS_SF_DENSE_PIECES=12 timeout 300 python bench.py --workload c2 --no-cpu --no-e2e --no-parity --steps 1 --warmup 1 2>/dev/null | grep ^SLT > gpurun_out/trace_slicef_ld8.txt
cp /tmp/product.so xpcs-eigen_b200/libxpcs_b200.so
wc -l gpurun_out/trace_slicef_ld4.txt gpurun_out/trace_slicef_ld8.txt
python profiles/trace_slice.py gpurun_out/trace_slicef_ld4.txt | head -70
