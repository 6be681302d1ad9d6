"""Derive the KITTI-seq00-shaped motion profile used by the synthetic streams.

Reads the reference's result artefact doc/tloam_00.txt (KITTI 3x4 camera-frame poses, one per line), converts
the frame-to-frame relative motions to the LiDAR axes (x_l = z_c, y_l = -x_c, z_l = -y_c) and stores them as
se(3) twists (upsilon, omega), float32, in tests/golden/motion_seq00.npy.  Runs only in the build container
(/root/reference does not exist on the GPU box); the .npy fixture is what travels.

usage: python tools/make_motion_profile.py [seq ...]
"""
import os
import sys

import numpy as np
from scipy.linalg import logm

REF = "/root/reference/doc"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
P = np.array([[0, 0, 1, 0], [-1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 0, 1]], dtype=np.float64)


def twist_of(T):
    L = np.real(logm(T))
    return np.array([L[0, 3], L[1, 3], L[2, 3], L[2, 1], L[0, 2], L[1, 0]])


def main(seqs):
    for seq in seqs:
        rows = np.loadtxt(os.path.join(REF, f"tloam_{seq}.txt"))
        poses = np.tile(np.eye(4), (rows.shape[0], 1, 1))
        poses[:, :3, :] = rows.reshape(-1, 3, 4)
        tw = []
        for k in range(1, poses.shape[0]):
            rel_c = np.linalg.inv(poses[k - 1]) @ poses[k]
            # re-orthonormalise the rotation (text round-off) before the log
            u, _, vt = np.linalg.svd(rel_c[:3, :3])
            rel_c[:3, :3] = u @ vt
            rel_l = P @ rel_c @ np.linalg.inv(P)
            tw.append(twist_of(rel_l))
        tw = np.asarray(tw, dtype=np.float32)
        path = os.path.join(OUT, f"motion_seq{seq}.npy")
        np.save(path, tw)
        print(path, tw.shape, "mean |v|", np.linalg.norm(tw[:, :3], axis=1).mean())


if __name__ == "__main__":
    main(sys.argv[1:] or ["00"])
