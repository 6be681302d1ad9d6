#!/bin/bash
# A/B: solver variants (single stream, inputs in HBM)
for v in default lie_inline before_ldlt default; do
  unset TLOAM_B200_LIB
  case $v in
    lie_inline) export TLOAM_B200_LIB=$PWD/build/variants/lie_inline.so ;;
    before_ldlt) export TLOAM_B200_LIB=$PWD/build/variants/before_ldlt.so ;;
  esac
  echo "== $v"
  python tools/multi_stream.py 1 2>&1 | tail -1
done
