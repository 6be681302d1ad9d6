#!/bin/bash
# ncu evidence for the segmentation front half ((f)-4): launch list + --set full of the two dominant kernels.
export PATH=/usr/local/cuda/bin:$PATH
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_seg_launches.csv python tools/segmentation_profile.py > gpurun_out/r2_seg_profile.log 2>&1
for k in k_os_seq k_ge_fit k_ee_section; do
  ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:$k -s 1 -c 1 -o gpurun_out/r2_$k python tools/segmentation_profile.py > gpurun_out/r2_$k.log 2>&1
  b=gpurun_out/r2_$k
  ncu -i $b.ncu-rep --page raw --csv > ${b}_raw.csv 2>/dev/null
  python tools/ncu_lines.py $b.ncu-rep 50 > ${b}_lines.txt 2>/dev/null
  python tools/ncu_lines.py $b.ncu-rep 40 smp > ${b}_lines_by_samples.txt 2>/dev/null
  rm -f $b.ncu-rep
done
