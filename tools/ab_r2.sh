#!/bin/bash
# single stream, inputs in HBM + in-kernel timers
for v in default default; do
  echo "== $v"
  python tools/multi_stream.py 1 2>&1 | tail -1
  python tools/timers.py 2>&1 | tail -5 | head -3
done
