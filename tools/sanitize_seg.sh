# compute-sanitizer passes over the segmentation front half ((f)-4) on a reduced scan, and over the chained device flow with
# the per-frame fitness on its side stream.      gpurun -- 'bash tools/sanitize_seg.sh'   (-> profiles/r2_compute_sanitizer.txt)
export TLOAM_B200_NO_GRAPH=1
cat > /tmp/seg_small.py <<PY
import sys; sys.path.insert(0, ".")
import numpy as np, tloam_b200
from tloam_b200 import synth
r = tloam_b200.LocalRegistration()
scan = synth.raw_scan(n_az=300)
ge = r.ground_extract(scan)
obj = np.ascontiguousarray(scan[ge["object"]]); beam = ge["beam"][ge["object"]].astype(np.float64)
os_ = r.object_segmentation(obj, min_seg=20)
sp = np.ascontiguousarray(obj[os_["segmented"]])
ee = r.extract_edge(sp, beam[os_["segmented"]], ring_min_num=16)
print("seg", len(scan), len(obj), len(sp), len(os_["sizes"]), len(ee["edge"]), len(ee["non_edge"]))
r.close()
PY
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool: segmentation chain"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/seg_small.py 2>&1 | tail -14
done
for tool in memcheck racecheck; do
  echo "== $tool: chained device flow test (fitness fork / join, scan prefetch)"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fitness or predicted" 2>&1 | tail -4
done
