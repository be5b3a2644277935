# round 2, session 4, the last 40 GPU-seconds: k_stream_chunk with the state and the event frames staged in shared memory
mkdir -p gpurun_out
timeout 28 python profiles/stream_check.py --tag occ5_smem --h 512 --w 512 --frames 16384 --occ 0.05 --chunks 2048,1024 --out gpurun_out/stream_occ5_r4c.json 2> gpurun_out/stream_occ5_r4c.err | cut -c1-250; echo "check exit $?"; tail -2 gpurun_out/stream_occ5_r4c.err
timeout 30 python -m pytest tests/test_gpu_stream.py -x -q --timeout 20 -k "equals_resident or event_walk or bright" > gpurun_out/pytest_r4c.log 2>&1; tail -3 gpurun_out/pytest_r4c.log
