#!/bin/bash
# tests + bench, summarised (run on the GPU box through gpurun)
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
tail -3 gpurun_out/bench_r2a.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2a.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["pageable_host"]["value"])
print(json.dumps(d.get("batched"))[:2500])
print(json.dumps(d["roofline"]["kernels"]))
PY
