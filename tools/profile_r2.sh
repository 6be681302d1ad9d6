#!/bin/bash
# ncu evidence for round 2 (run under gpurun on ONE GPU); outputs under gpurun_out/
set -x
export PATH=/usr/local/cuda/bin:$PATH
# (1) launch list of the single-stream frame path (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/r2_launches.csv \
    python tools/profile_frame.py 8 > gpurun_out/r2_launches.log 2>&1
# (2) --set full of the frame kernels (3 launches each)
for k in k_correspond k_begin_frame k_map_insert k_map_scatter k_map_offsets k_map_bbox k_stage_source; do
  TLOAM_B200_FINE=0 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 2 -o gpurun_out/r2_$k python tools/profile_frame.py > gpurun_out/r2_$k.log 2>&1
done
# the evaluation kernels: launch 0 is k_eval<first>, launches 1.. are ACTIVE k_eval<false> (solver tail included);
# dense PC sampling (every 64 cycles) so that the single-warp solver tail shows up by source line
TLOAM_B200_FINE=0 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:k_eval -s 0 -c 3 -o gpurun_out/r2_k_eval python tools/profile_frame.py > gpurun_out/r2_k_eval.log 2>&1
# (2b) two-level grid on config 3 (full size): the second-level build and the search
ncu --set full --clock-control none --import-source on -k regex:k_correspond_fine -s 1 -c 1 -o gpurun_out/r2_fine python tools/config3.py 1.0 > gpurun_out/r2_fine.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_map_fine -s 1 -c 2 -o gpurun_out/r2_map_fine python tools/config3.py 1.0 > gpurun_out/r2_map_fine.log 2>&1
# (3) dense path (config 3, reduced: the kernel is the same)
TLOAM_B200_DENSE=1 ncu --set full --clock-control none --import-source on -k regex:k_correspond_dense -s 1 -c 2 -o gpurun_out/r2_dense python tools/config3.py 0.25 > gpurun_out/r2_dense.log 2>&1
TLOAM_B200_DENSE=1 ncu --set full --clock-control none --import-source on -k regex:k_qbin -s 3 -c 3 -o gpurun_out/r2_qbin python tools/config3.py 0.25 > gpurun_out/r2_qbin.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw metric tables and the per-source-line summaries, drop the reports
for r in gpurun_out/r2_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
  python tools/ncu_lines.py $r 60 > ${b}_lines.txt 2>/dev/null
  python tools/ncu_lines.py $r 40 smp > ${b}_lines_by_samples.txt 2>/dev/null
  rm -f $r
done
ls -la gpurun_out/ | head -60
