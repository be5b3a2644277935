# round 2, session 4, first shot: the online multi-tau on a GPU -- its tests, then one streamed C3 and one 5 % job against the resident path
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_stream.py -q --timeout 60 > gpurun_out/pytest_r4a.log 2>&1; tail -25 gpurun_out/pytest_r4a.log
timeout 80 python profiles/stream_check.py --tag c3 --out gpurun_out/stream_c3_r4a.json 2> gpurun_out/stream_c3_r4a.err; echo "stream c3 exit $?"; tail -3 gpurun_out/stream_c3_r4a.err
timeout 60 python profiles/stream_check.py --tag occ5 --h 512 --w 512 --frames 16384 --occ 0.05 --chunk 2048 --out gpurun_out/stream_occ5_r4a.json 2> gpurun_out/stream_occ5_r4a.err; echo "stream occ5 exit $?"; tail -3 gpurun_out/stream_occ5_r4a.err
