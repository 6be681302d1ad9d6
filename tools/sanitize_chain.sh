export TLOAM_B200_NO_GRAPH=1
cat > /tmp/seg_one.py <<PY
import sys; sys.path.insert(0, ".")
import numpy as np, tloam_b200
from tloam_b200 import synth
r = tloam_b200.LocalRegistration()
scan = synth.raw_scan(n_az=300)
o = r.segment_scan(scan, ring_min_num=16, dcvc=dict(min_seg=20))
print("segment_scan", {k: len(v) for k, v in o.items()})
big = np.random.default_rng(0).uniform(-40, 40, (200000, 3)) * np.array([1, 1, 0.05])
r.set_input_target([big[:5000].copy(), big[:20000].copy(), big, big[:90000].copy()])       # pageable: staged by the library
print("knn", r.knn(2, big[:100], 0.3, 5)[2].sum())
r.close()
PY
for tool in memcheck racecheck; do
  echo "== $tool: tloam_b200_segment_scan + pageable staging (set_target of a 4.8 MB cloud)"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/seg_one.py 2>&1 | tail -4
done
