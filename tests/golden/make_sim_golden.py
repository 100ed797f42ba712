"""Generate tests/golden/sim_update.npz from the REFERENCE'S OWN CODE (simulator-form agent update, MAIN:1408-1417).

Run in the build container only (needs /root/reference):  python tests/golden/make_sim_golden.py
Drives oracle/_ref/liblqro_ref.so through the canonical 6-quad scenario (MAIN:1163-1218, 210 steps) and stores
  * the reference's globals + gains in the layout of lqro_sim_model (458 doubles),
  * for a set of recorded steps: the full per-agent state before the update (x, Rot, P, xTrue, RotTrue, pGoal),
    newV, the 22 standard normals per agent the reference drew (MAIN:450, 1413) and the full state after it,
  * the MSVC rand() state at the start of every step and the whole trajectory (x, xTrue, newV per step),
so that the GPU parity tests run where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from _libs import Reference, RefSim  # noqa: E402
from make_golden import hexagon  # noqa: E402

from lqr_obstacles_b200.api import msvc_normals  # noqa: E402  (host-only helper of the product library)

STEPS = 210
RECORD = [0, 1, 5, 30, 60, 75, 90, 105, 120, 150, 209]
KEYS = ("x", "Rot", "P", "xTrue", "RotTrue")


def main():
    ref = Reference()
    starts, goals = hexagon()
    sim = RefSim(ref, starts, goals, seed=1)
    n = sim.n
    out = {"model": sim.model(), "starts": starts, "goals": goals, "record_steps": np.array(RECORD)}
    st0 = sim.full_state()
    for k in KEYS + ("pGoal", "vGoal"):
        out["init_" + k] = st0[k]
    seeds, traj_x, traj_xt, traj_newv = [], [], [], []
    for t in range(STEPS):
        newV = sim.avoid(H=30)[0]
        before = sim.full_state()
        seed = sim.seed()
        seeds.append(seed)
        normals, seed_after = msvc_normals(seed, 22 * n)
        sim.update()
        assert sim.seed() == seed_after, "lqro_msvc_normals is out of step with the reference's normal()"
        after = sim.full_state()
        traj_x.append(after["x"]); traj_xt.append(after["xTrue"]); traj_newv.append(newV)
        if t in RECORD:
            for k in KEYS + ("pGoal",):
                out["s%d_before_%s" % (t, k)] = before[k]
            out["s%d_newV" % t] = newV
            out["s%d_normals" % t] = normals.reshape(n, 22)
            for k in KEYS + ("vGoal",):
                out["s%d_after_%s" % (t, k)] = after[k]
    out["seeds"] = np.array(seeds, dtype=np.uint32)
    out["traj_x"] = np.array(traj_x)
    out["traj_xTrue"] = np.array(traj_xt)
    out["traj_newV"] = np.array(traj_newv)
    sim.close()
    path = os.path.join(HERE, "sim_update.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
