# usage: bash tools/profile.sh v5   (inside gpurun) -- refreshes every artefact kept under profiles/
V=$1
export PYTHONUNBUFFERED=1
TLOAM_B200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1$V.csv python tools/profile_frame.py 3 > gpurun_out/ncu_a.log 2>&1
TLOAM_B200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_correspond|k_eval" -s 12 -c 12 -o gpurun_out/prof_r1$V -f python tools/profile_frame.py 2 > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/prof_r1$V.ncu-rep --page raw --csv > gpurun_out/raw_r1$V.csv 2>/dev/null
timeout 300 python bench.py --steps 30 --warmup 3 2>gpurun_out/bench_$V.err > gpurun_out/bench_$V.json
timeout 400 python bench.py --impl reference --steps 6 --warmup 1 2>gpurun_out/bench_ref_$V.err > gpurun_out/bench_ref_$V.json
timeout 300 python tools/config3.py 1.0 > gpurun_out/config3_$V.json 2>gpurun_out/config3_$V.err
timeout 300 python tools/multi_stream.py 1 2 4 8 > gpurun_out/multi_stream_$V.json 2>gpurun_out/multi_stream_$V.err
timeout 300 python tools/multi_stream.py --e2e 1 2 4 8 >> gpurun_out/multi_stream_$V.json 2>>gpurun_out/multi_stream_$V.err
tail -c 600 gpurun_out/bench_$V.json; tail -c 400 gpurun_out/bench_ref_$V.json; tail -3 gpurun_out/multi_stream_$V.json
