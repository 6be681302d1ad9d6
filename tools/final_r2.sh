#!/bin/bash
# Final evidence pass of round 2 (one B200, under gpurun): tests, sanitizer, launch lists, ncu --set full of the new kernels,
# the bench lines.  Everything lands in gpurun_out/ with r2_ names; the summaries are copied to profiles/ afterwards.
export PATH=/usr/local/cuda/bin:$PATH
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_gpu_tests.txt
bash tools/sanitize_seg.sh > gpurun_out/r2_compute_sanitizer.txt 2>&1
bash tools/profile_seg.sh
python tools/segmentation_bench.py 1875 2>&1 | tail -1 > gpurun_out/r2_segmentation_bench.json
python tools/feature_bench.py 50000 120000 2>&1 | tail -2 > gpurun_out/r2_feature_extraction.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_feature_launches.csv python tools/feature_profile.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_chain_launches.csv python tools/profile_chain.py 6 > gpurun_out/r2_chain.log 2>&1
python tools/e2e_probe.py 23 > gpurun_out/r2_e2e_probe.txt 2>&1
python tools/e2e_probe.py 23 pageable >> gpurun_out/r2_e2e_probe.txt 2>&1
TLOAM_B200_NO_HOST_STAGE=1 python tools/e2e_probe.py 23 pageable >> gpurun_out/r2_e2e_probe.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r2_bench_config3.json 2>> gpurun_out/r2_bench.err
tail -2 gpurun_out/r2_gpu_tests.txt; grep -c "ERROR SUMMARY: 0 errors\|0 hazards" gpurun_out/r2_compute_sanitizer.txt
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'pageable',d['e2e']['pageable_host']['value'],'chain',d['stream_device_submap']['value'],'batched',d['batched']['value'],d['batched']['groups']['value'])
print('seg',d['segmentation']['gpu_ms_per_call'],d['segmentation'].get('cpu_port_ms_per_call'))
print('feat',d['feature_extraction']['gpu_ms_per_call'])
"
