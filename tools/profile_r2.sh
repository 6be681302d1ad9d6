#!/bin/bash
# ncu evidence for round 2 (run under gpurun on ONE GPU); outputs under gpurun_out/
set -x
export PATH=/usr/local/cuda/bin:$PATH
# (1) launch list of the single-stream frame path (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/r2_launches.csv \
    python tools/profile_frame.py 8 > gpurun_out/r2_launches.log 2>&1
# (2) --set full of the frame kernels (3 launches each)
for k in k_first k_eval k_begin_frame k_map_insert k_map_scatter k_map_offsets k_map_bbox k_stage_source; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 2 -o gpurun_out/r2_$k python tools/profile_frame.py > gpurun_out/r2_$k.log 2>&1
done
# (3) dense path (config 3, reduced: the kernel is the same)
TLOAM_B200_DENSE=1 ncu --set full --clock-control none --import-source on -k regex:k_correspond_dense -s 1 -c 2 -o gpurun_out/r2_dense python tools/config3.py 0.25 > gpurun_out/r2_dense.log 2>&1
TLOAM_B200_DENSE=1 ncu --set full --clock-control none --import-source on -k regex:k_qbin -s 3 -c 3 -o gpurun_out/r2_qbin python tools/config3.py 0.25 > gpurun_out/r2_qbin.log 2>&1
ls -la gpurun_out/*.ncu-rep
