import sys; sys.path.insert(0, ".")
import tloam_b200
from tloam_b200 import synth
r = tloam_b200.LocalRegistration()
p = synth.general_cloud(50000, seed=77)
for _ in range(2): out = r.extract_planar_sphere(p)
r.close()
