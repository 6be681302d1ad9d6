run() { timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/ab_$1.err > gpurun_out/ab_$1.json; python - <<PY
import json
d=json.load(open("gpurun_out/ab_$1.json"))
print("$1", round(d["value"],1), round(d["e2e"]["value"],1), round(d["stream_device_submap"]["value"],1), {n:round(v["avg_us"],1) for n,v in d["roofline"]["kernels"].items()}, d.get("feature_extraction"))
PY
}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run inplace
