"""Generates tests/golden/tmain_keyframes.npy: the key-frame trajectory (position of every quadrotor after every time step,
as handed to Callisto, FP32) of the UNMODIFIED reference main() -- 6 quadrotors, 210 steps, MSVC rand() seed 1 -- run through
oracle/_ref (ref_run_tmain).  Build container only (needs /root/reference).  usage: python tests/golden/make_tmain_golden.py"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _libs import Reference  # noqa: E402

R = Reference()
lib = R.lib
lib.lqr_ref_clear_keyframes()
devnull = os.open(os.devnull, os.O_RDONLY)
os.dup2(devnull, 0)          # the reference ends with `std::cin >> k` (MAIN:1421-1422)
saved = os.dup(1)
os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
lib.ref_run_tmain()
os.dup2(saved, 1)
n = lib.lqr_ref_num_keyframes()
g, t, p = ctypes.c_int(), ctypes.c_float(), (ctypes.c_float * 3)()
rows = []
for i in range(n):
    lib.lqr_ref_get_keyframe(i, ctypes.byref(g), ctypes.byref(t), p)
    rows.append((g.value, t.value, p[0], p[1], p[2]))
rows = np.array(rows, dtype=np.float64)
groups = sorted(np.unique(rows[:, 0]))[0::2]    # cloned in the order quadrotor, ellipse per quad
traj = np.stack([rows[rows[:, 0] == gid][:, 2:5] for gid in groups], axis=1).astype(np.float32)   # (steps, quads, 3)
np.save(os.path.join(HERE, "tmain_keyframes.npy"), traj)
print("wrote", traj.shape, "key frames;", n, "records")
