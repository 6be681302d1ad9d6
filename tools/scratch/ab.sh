run() { timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/ab_$1.json; python - <<PY
import json
d=json.load(open("gpurun_out/ab_$1.json"))
print("$1", round(d["value"],1), round(d["e2e"]["value"],1), {n:round(v["avg_us"],1) for n,v in d["roofline"]["kernels"].items() if n in ("correspond","eval_first","eval")})
PY
}
run base
cp tloam_b200/libtloam_b200.so /tmp/keep.so
for v in mb6 mb8; do cp tloam_b200/libtloam_b200_$v.so tloam_b200/libtloam_b200.so; run $v; done
