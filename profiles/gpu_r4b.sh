# round 2, session 4, last shot: whole GPU suite (with the stream tests), then the online multi-tau at 5 % occupancy:
# chunk lengths, event walk on / off, host-buffer path, one ncu --set full capture of k_stream_chunk, C5 geometry at 5 %
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_r4b.log 2>&1; tail -4 gpurun_out/pytest_r4b.log
O5="--h 512 --w 512 --frames 16384 --occ 0.05"
timeout 70 python profiles/stream_check.py --tag occ5 $O5 --chunks 1024,2048,4096 --also-bins-only --host --out gpurun_out/stream_occ5_r4b.json 2> gpurun_out/stream_occ5_r4b.err | cut -c1-300; echo "sweep exit $?"; tail -2 gpurun_out/stream_occ5_r4b.err
timeout 60 ncu --set full --clock-control none --import-source on -k regex:'k_stream_chunk' -s 2 -c 1 -o gpurun_out/prof_stream_r04 -f python profiles/stream_check.py $O5 --chunk 2048 --profile > gpurun_out/ncu_stream_r04.log 2>&1; tail -1 gpurun_out/ncu_stream_r04.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/prof_stream_r04.ncu-rep > gpurun_out/ncu_stream_r04.txt 2>/dev/null; head -28 gpurun_out/ncu_stream_r04.txt
timeout 70 python profiles/stream_big.py --out gpurun_out/stream_c5_occ5_r4b.json 2> gpurun_out/stream_c5_occ5_r4b.err | cut -c1-1200; echo "big exit $?"; tail -3 gpurun_out/stream_c5_occ5_r4b.err
