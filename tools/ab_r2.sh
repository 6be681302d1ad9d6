#!/bin/bash
# A/B: fused vs un-fused first evaluation, k_first occupancy variants (single stream and batched)
for v in default nofuse mb6 mb8; do
  unset TLOAM_B200_NO_FUSE TLOAM_B200_LIB
  case $v in
    nofuse) export TLOAM_B200_NO_FUSE=1 ;;
    mb6) export TLOAM_B200_LIB=$PWD/build/variants/first_mb6.so ;;
    mb8) export TLOAM_B200_LIB=$PWD/build/variants/first_mb8.so ;;
  esac
  echo "== $v"
  python tools/multi_stream.py 1 2>&1 | tail -1
  python tools/batch_bench.py 8 16 2>&1 | tail -2
done
