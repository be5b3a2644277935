mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2x.log 2>&1; tail -4 gpurun_out/pytest_r2x.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_multigpu.py::test_slab_partition_kernels_on_one_gpu" -m gpu -q > gpurun_out/sanitize_runs.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_runs.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_multigpu.py::test_slab_partition_kernels_on_one_gpu" -m gpu -q > gpurun_out/sanitize_runs_race.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_runs_race.log
timeout 300 python bench.py --no-cpu --steps 5 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); k=j['kernels']; print('c3 ms/step %.3f e2e %.3f parity %s'%(j['ms_per_step'], j['e2e']['ms_per_step'], j['parity']['ok']), {x:round(k[x]['ms_per_step'],3) for x in ('k_multitau_warp','k_finalize','k_scatter_rec','k_hist')})"
