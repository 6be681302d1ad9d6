#!/bin/bash
# A/B: solver variants (single stream, inputs in HBM)
for v in default unrolled default unrolled; do
  unset TLOAM_B200_LIB
  case $v in
    unrolled) export TLOAM_B200_LIB=$PWD/build/variants/unrolled.so ;;
  esac
  echo "== $v"
  python tools/multi_stream.py 1 2>&1 | tail -1
  python tools/timers.py 2>&1 | tail -4 | head -1
  python tools/timers.py 2>&1 | tail -2 | head -1
done
