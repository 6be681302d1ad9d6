#!/bin/bash
# A/B: k_eval occupancy variants in batched mode
for v in default emb4 emb5; do
  unset TLOAM_B200_LIB
  case $v in
    emb4) export TLOAM_B200_LIB=$PWD/build/variants/eval_mb4.so ;;
    emb5) export TLOAM_B200_LIB=$PWD/build/variants/eval_mb5.so ;;
  esac
  echo "== $v"
  python tools/batch_bench.py 8 2>&1 | tail -1
done
