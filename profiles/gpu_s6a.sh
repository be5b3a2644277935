# session 6 start: sanity suite on the rebuilt library, source-level ncu of the multi-tau kernels (c3) and the dense path (c2)
TAG=${1:-s6a}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tail -9
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_multitau_warp' -c 1 -o gpurun_out/prof_mtw_$TAG -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_mtw.log 2>&1
tail -2 gpurun_out/ncu_mtw.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_dense_filter|k_multitau|k_finalize|k_scatter' -c 4 -o gpurun_out/prof_c2_$TAG -f python bench.py --workload c2 --frames 4000 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log
ls -la gpurun_out
