# compute-sanitizer passes over a reduced workload (scale 0.05: F = 2k, M = 25k) and a small feature extraction.
#   gpurun -- 'bash tools/sanitize.sh'        (results: profiles/r1_v5_compute_sanitizer.txt)
export TLOAM_B200_NO_GRAPH=1
cat > /tmp/fe_small.py <<PY
import sys; sys.path.insert(0, ".")
import numpy as np, tloam_b200
from tloam_b200 import synth
r = tloam_b200.LocalRegistration()
p = synth.general_cloud(4000, seed=3)
out = r.extract_planar_sphere(p)
print("fe lists", [len(x) for x in out])
a = np.random.default_rng(0).uniform(-5, 5, (3000, 3))
print("vox", r.voxel_down_sample(a, 0.5).shape)
r.close()
PY
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool: frame"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/profile_frame.py 2 0.05 2>&1 | grep -v "^frame\|^$" | tail -6
  echo "== $tool: feature extraction + voxel"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/fe_small.py 2>&1 | tail -5
done
