"""The CPU restatement against the reference's OWN code compiled from /root/reference
(oracle/_ref/liblqro_ref.so).  Skipped where neither the prebuilt library nor the tree exists."""
import numpy as np
import pytest

from conftest import state_pair_random


def test_pairs_random_bit_exact(oracle, reference, golden):
    A, B, L, E = golden["dyn_A"], golden["dyn_B"], golden["dyn_L"], golden["dyn_E"]
    cm = reference.control_matrices()
    assert np.array_equal(cm["A"], A) and np.array_equal(cm["L"], L)  # the golden file is this build's output
    sph = oracle.spheres()
    rng = np.random.default_rng(11)
    u = reference.nominal_input()
    counts = {0: 0, 1: 0, 2: 0}
    for H in (30, 20):
        NF, T, _ = oracle.horizon(A, B, L, E, H)
        for t in range(150):
            xi, xj = state_pair_random(rng, u, t % 3)
            pn = rng.normal(size=3)
            r = reference.pair(A, B, L, E, xi, xj, H, 30.0, prev_normal=pn)
            o = oracle.pair(NF, T, sph, xi, xj, 30.0, fix_facet0=0, prev_normal=pn)
            q = oracle.candidates(NF, T, sph, xi, xj)
            assert np.array_equal(q, r["ellipsoids"])                          # createObstacle
            assert np.array_equal(oracle.reachable(q, xi, xj), r["reach_idx"])  # findReachableObstacle (pow folding)
            assert o["status"] == r["status"]
            counts[r["status"]] += 1
            if r["status"]:
                assert np.array_equal(o["gjk_stats"], r["gjk_stats"])
                assert o["distance"] == r["distance"]
                assert np.array_equal(o["plane"], r["plane"])
    assert counts[1] > 50 and counts[2] > 10


def test_gjk_random_clouds(oracle, reference):
    rng = np.random.default_rng(5)
    for t in range(200):
        n = int(rng.integers(5, 400))
        pts = rng.normal(size=(n, 3)) * rng.uniform(0.5, 20) + rng.normal(size=3) * 5
        vrel = rng.normal(size=3) * rng.uniform(0.1, 30)
        d, nrm, st = reference.run_gjk(vrel, pts)
        d2, w1, w2, st2 = oracle.gjk(vrel, pts)
        assert np.sqrt(d2) == d and np.array_equal(st, st2)


def test_lp_random_bit_exact(oracle, reference):
    rng = np.random.default_rng(7)
    for t in range(1500):
        n = int(rng.integers(0, 64))
        normals = rng.normal(size=(n, 3))
        normals /= np.linalg.norm(normals, axis=1)[:, None]
        pts = rng.normal(size=(n, 3)) * rng.uniform(0.1, 5)
        if t % 3 == 0:
            pts = -normals * rng.uniform(0, 3, size=(n, 1))
        if t % 5 == 0 and n > 3:
            normals[2] = -normals[0]
            pts[2] = pts[0] + normals[0] * rng.uniform(0.1, 2)
        planes = np.concatenate([pts, normals], axis=1).astype(np.float32)
        goal = rng.normal(size=3) * rng.uniform(0.1, 150)
        assert np.array_equal(reference.calculate_new_v(planes, goal), oracle.new_velocity(planes, goal))


def test_hull_order_matches_qhull(oracle):
    """oracle/hull_cpu.cpp against scipy's qhull_r 8.0.2: same facets, same order, same vertex order,
    bit-identical planes, on 6-digit clouds of the kind the path produces."""
    sp = pytest.importorskip("scipy.spatial")
    rng = np.random.default_rng(0)
    nexact = ntotal = 0
    for t in range(120):
        n = int(rng.integers(8, 1500))
        kind = t % 4
        if kind == 0:
            p = rng.normal(size=(n, 3)) * [10, 10, 20]
        elif kind == 1:
            p = rng.normal(size=(n, 3))
            p /= np.linalg.norm(p, axis=1)[:, None]
            p = p * [8, 9, 20] + rng.normal(size=(n, 3)) * 0.5
        elif kind == 2:
            p = rng.uniform(-30, 30, size=(n, 3))
        else:  # union of ellipsoid shells, like an LQR-obstacle
            k = rng.integers(0, 12, size=n)
            d = rng.normal(size=(n, 3))
            d /= np.linalg.norm(d, axis=1)[:, None]
            p = d * (np.array([0.5, 0.5, 1.5]) * (1 + k[:, None] * 0.7)) + k[:, None] * np.array([1.5, 0.3, 0.1])
        p = oracle.round6(p)
        planes, verts, order, info = oracle.hull(p)
        h = sp.ConvexHull(p)
        assert info["degenerate"] == 0
        assert np.array_equal(verts, h.simplices), (t, n, kind)
        # planes are bit-identical except on sliver facets, where qhull re-derives the normal by Gaussian
        # elimination (qh_sethyperplane_gauss after `nearzero`): ~1e-14, five orders below the 1e-9 bar
        diff = np.abs(planes - h.equations).max(axis=1)
        assert diff.max() < 1e-12
        nexact += int((diff == 0).sum())
        ntotal += len(diff)
    assert nexact > 0.99 * ntotal


def test_reference_hull_branch_same_with_scipy_qhull(oracle, reference, golden):
    """Run the reference's convexHull through its text-file round trip with scipy/qhull as the qconvex
    stand-in and with oracle/hull_cpu.cpp: identical distance and normal."""
    pytest.importorskip("scipy.spatial")
    rng = np.random.default_rng(2)
    for t in range(25):
        n = int(rng.integers(50, 900))
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1)[:, None]
        pts = d * [6, 7, 15] * rng.uniform(0.6, 1.0, size=(n, 1))
        vrel = rng.normal(size=3)
        pn = rng.normal(size=3)
        reference.use_scipy_hull(True)
        d1, n1 = reference.convex_hull(vrel, pts, prev_normal=pn)
        reference.use_scipy_hull(False)
        d2, n2 = reference.convex_hull(vrel, pts, prev_normal=pn)
        nf, d3, n3, best, deg = oracle.hull_escape(pts, vrel, fix_facet0=0, prev_normal=pn)
        assert d1 == d2 == d3 and np.array_equal(n1, n2) and np.array_equal(n2, n3)


def test_six_quad_trajectory_reproduces_unmodified_main(reference):
    """The harness loop (ref_sim_*) must be the reference's own main loop: compare the key-frame
    positions the UNMODIFIED _tmain emits (through the stubbed CAL_AddGroupKeyState) with the harness
    run of the same scenario, first 40 steps."""
    import ctypes
    import os
    import subprocess
    import sys

    # run in a child process: _tmain reads std::cin at the end and uses the global RNG state
    code = r'''
import sys, ctypes, numpy as np, os
sys.path.insert(0, %r)
from _libs import Reference, RefSim
from lqr_obstacles_b200.swarm import hexagon6
R = Reference()
lib = R.lib
S, G = hexagon6()
sim = RefSim(R, S, G, seed=1)
traj = []
for t in range(40):
    newV = sim.avoid(30, 30.0)[0]
    sim.update()
    traj.append(sim.state()[1][:, :3].copy())
sim.close()
np.save(os.path.join(os.environ["LQRO_OUT"], "harness.npy"), np.array(traj))
''' % os.path.dirname(os.path.abspath(__file__))
    # the unmodified main runs 210 steps (26 s with a process spawn per hull; ~3 s in-process)
    code2 = r'''
import sys, ctypes, numpy as np, os
sys.path.insert(0, %r)
from _libs import Reference
R = Reference()
lib = R.lib
lib.lqr_ref_clear_keyframes()
devnull = os.open(os.devnull, os.O_RDONLY); os.dup2(devnull, 0)
out = os.open(os.devnull, os.O_WRONLY); os.dup2(out, 1)
lib.ref_run_tmain()
n = lib.lqr_ref_num_keyframes()
g, t, p = ctypes.c_int(), ctypes.c_float(), (ctypes.c_float * 3)()
rows = []
for i in range(n):
    lib.lqr_ref_get_keyframe(i, ctypes.byref(g), ctypes.byref(t), p)
    rows.append((g.value, t.value, p[0], p[1], p[2]))
np.save(os.path.join(os.environ["LQRO_OUT"], "tmain.npy"), np.array(rows))
''' % os.path.dirname(os.path.abspath(__file__))
    import tempfile

    d = tempfile.mkdtemp()
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), LQRO_OUT=d)
    subprocess.check_call([sys.executable, "-c", code], cwd=d, env=env)
    subprocess.check_call([sys.executable, "-c", code2], cwd=d, env=env)
    harness = np.load(os.path.join(d, "harness.npy"))            # (40, 6, 3) xTrue positions after each update
    rows = np.load(os.path.join(d, "tmain.npy"))                 # key frames: 2 per quad per step (quad, ellipse)
    groups = np.unique(rows[:, 0])
    assert len(groups) == 12
    quad_groups = sorted(groups)[0::2]                            # cloned in the order quadrotor, ellipse per quad
    for qi, gid in enumerate(quad_groups):
        pos = rows[rows[:, 0] == gid][:, 2:5]
        assert len(pos) == 210
        assert np.array_equal(pos[:40], harness[:, qi, :].astype(np.float32)), qi
