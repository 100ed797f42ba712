"""CPU check of the LOGIC of stage 2 of the pair kernel: the kernel sources csrc/lqro_pair.cuh (pair_translate_steps,
pair_scan_cloud) and csrc/lqro_gjk.cuh (gjk_warp_lanes, gjk_warp) compiled as plain C++ on the 32-thread lockstep emulation of the
warp primitives (tests/hull_host/warp_emul.h) against the oracle's restatement of createObstacle / findReachableObstacle
(MAIN:683-706) and of Cameron's gjk_distance as run_gjk calls it (MAIN:712-746, GJK.c:296-501).  Everything is compared bit for
bit: the reachable cloud (points and order), d^2, both witness points and the number of G-tests.  The GPU tests
(tests/test_gpu_parity.py, tests/test_gpu_headline.py) run the same source on the device."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import state_pair_random

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pair_host", "pair_host.cpp")
SO = os.path.join(HERE, "pair_host", "libpairhost.so")
CSRC = os.path.join(HERE, "..", "lqr_obstacles_b200", "csrc")
DEPS = [SRC, os.path.join(HERE, "hull_host", "warp_emul.h")] + [os.path.join(CSRC, f) for f in ("lqro_common.cuh", "lqro_gjk.cuh", "lqro_pair.cuh")]
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


# builds of the emulation: the shipped configuration; every written-out default step on (incl. the four-point one) + a decision band
# of the step-major candidate scan so wide that its redo path runs on most pairs; the general lane-parallel step and the
# candidate-major scan alone
BUILDS = {"shipped": [], "all_steps": ["-DLQRO_GJK_STEP3=1", "-DLQRO_GJK_STEP4=1", "-DLQRO_SCAN_TEST_WIDE_BAND=1"], "general": ["-DLQRO_GJK_STEP2=0", "-DLQRO_GJK_STEP3=0", "-DLQRO_GJK_STEP4=0", "-DLQRO_SCAN_STEPMAJOR=0"]}


class PairHost:
    def __init__(self, build="shipped"):
        so = SO.replace(".so", "_%s.so" % build)
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in DEPS):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++20", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared", "-pthread"] + BUILDS[build]
                                  + ["-o", so, SRC])
        self.lib = lib = ctypes.CDLL(so)
        lib.pairhost_stage2.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5 + [ctypes.c_double, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.pairhost_gjk.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

    def stage2(self, NF, T, sph, xi, xj, templ, kmask=None, cap=256, lanes=1, max_speed=30.0):
        H, Pn = NF.shape[0], sph.shape[0]
        NFt = np.ascontiguousarray(NF.reshape(3 * H, 16).T)  # [16][3H], NFt[m][3k+c] = NF[k][c][m]
        cloud, g = np.zeros((H * Pn, 3)), np.zeros(8)
        km = (1 << H) - 1 if kmask is None else kmask
        n = self.lib.pairhost_stage2(H, Pn, templ, P(NFt), P(np.ascontiguousarray(T)), P(np.ascontiguousarray(sph)), P(xi), P(xj), max_speed, km, cap,
                                     lanes, P(cloud), P(g))
        return cloud[:n], g

    def gjk(self, vrel, pts, cap=256, lanes=1):
        g = np.zeros(8)
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        self.lib.pairhost_gjk(P(np.ascontiguousarray(vrel, dtype=np.float64)), P(pts), len(pts), cap, lanes, P(g))
        return g


@pytest.fixture(scope="module", params=list(BUILDS))
def host(request):
    return PairHost(request.param)


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _same_gjk(g, want, n):
    d2, w1, w2, st = want
    return (_bits(g[0]) == _bits(d2)).all() and (_bits(g[1:4]) == _bits(w1)).all() and (_bits(g[4:7]) == _bits(w2)).all() and int(g[7]) == int(st[0])


def _tables(oracle, golden, H):
    NF, T, _ = oracle.horizon(golden["dyn_A"], golden["dyn_B"], golden["dyn_L"], golden["dyn_E"], H)
    return NF, T, oracle.spheres(50)


@pytest.mark.parametrize("H", [20, 30])
def test_stage2_random_pairs(host, oracle, golden, H):
    """cloud + GJK of random pairs at all three distance regimes (inside / near / far), both walks of the candidate scan, the
    shared-memory cap of the kernel and a small one (overflow arrays), lane-parallel and scalar GJK"""
    NF, T, sph = _tables(oracle, golden, H)
    rng = np.random.default_rng(11 + H)
    u = golden["dyn_u"] if "dyn_u" in golden.files else np.full(4, 1.2)
    seen = dict(empty=0, outside=0, inside=0)
    for t in range(36):
        xi, xj = state_pair_random(rng, u, (2, 2, 1)[t % 3])
        if t % 4 == 0:  # same attitude / rotor state: the zero-term skip of Translate_k
            xj[6:] = xi[6:]
        q = oracle.candidates(NF, T, sph, xi, xj)
        idx = oracle.reachable(q, xi, xj)
        want_cloud = q[idx]
        vrel = xi[3:6] - xj[3:6]
        want = oracle.gjk(vrel, want_cloud) if len(idx) > 4 else None
        for templ, cap, lanes in ((1, 256, 1), (0, 256, 1), (1, 40, 1), (0, 40, 0), (1, 256, 0)):
            cloud, g = host.stage2(NF, T, sph, xi, xj, templ, cap=cap, lanes=lanes)
            assert len(cloud) == len(idx), (t, templ, cap)
            assert (_bits(cloud) == _bits(want_cloud)).all(), (t, templ, cap)
            if want is not None:
                assert _same_gjk(g, want, len(idx)), (t, templ, cap, lanes, g, want)
        if want is None:
            seen["empty"] += 1
        elif want[0] < 1e-8:
            seen["inside"] += 1
        else:
            seen["outside"] += 1
    assert seen["outside"] >= 5 and seen["inside"] >= 2, seen


def test_stage2_step_subsets(host, oracle, golden):
    """a mask of surviving steps that still holds every step with a reachable point gives the same cloud as all steps"""
    H = 20
    NF, T, sph = _tables(oracle, golden, H)
    rng = np.random.default_rng(5)
    u = np.full(4, 1.2)
    for t in range(12):
        xi, xj = state_pair_random(rng, u, 1)
        q = oracle.candidates(NF, T, sph, xi, xj)
        idx = oracle.reachable(q, xi, xj)
        need = 0
        for k in set((idx // 50).tolist()):
            need |= 1 << k
        extra = int(rng.integers(0, 1 << H))
        for templ in (0, 1):
            cloud, _ = host.stage2(NF, T, sph, xi, xj, templ, kmask=need | extra)
            assert (_bits(cloud) == _bits(q[idx])).all()
            cloud, _ = host.stage2(NF, T, sph, xi, xj, templ, kmask=need)
            assert (_bits(cloud) == _bits(q[idx])).all()


def test_gjk_random_clouds(host, oracle):
    """GJK alone on generic clouds: point well outside, near the surface, inside; tiny and large clouds; degenerate (coplanar,
    collinear, duplicate) ones that drive the simplex through the size-2 / size-3 reductions and the backup procedure"""
    rng = np.random.default_rng(3)
    n_iters = {}
    for t in range(160):
        n = int(rng.choice([5, 6, 9, 17, 33, 64, 150, 300, 700]))
        kind = t % 8
        pts = rng.normal(size=(n, 3)) * rng.uniform(0.2, 3.0, size=3)
        if kind == 5:
            pts[:, 2] = 0.25  # coplanar
        if kind == 6:
            pts = np.outer(rng.normal(size=n), rng.normal(size=3)) + 0.5  # collinear
        if kind == 7:
            pts[n // 2:] = pts[: n - n // 2]  # duplicates
        c = pts.mean(axis=0)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        vrel = c + d * (0.0, 0.3, 1.0, 2.5, 8.0, 3.0, 2.0, 1.5)[kind] * np.abs(pts - c).max()
        if kind == 1 and t % 16 == 1:
            vrel = pts[int(rng.integers(n))].copy()  # exactly on a vertex
        want = oracle.gjk(vrel, pts)
        for cap, lanes in ((256, 1), (32, 1), (256, 0)):
            g = host.gjk(vrel, pts, cap=cap, lanes=lanes)
            assert _same_gjk(g, want, n), (t, n, kind, cap, lanes, g, want)
        n_iters[int(g[7])] = n_iters.get(int(g[7]), 0) + 1
    assert len(n_iters) >= 4, n_iters


def test_gjk_backup_procedure(host, oracle):
    """clouds far from the origin (cancellation makes the delta signs inconsistent): no subset passes the default procedure and
    Cameron's backup procedure (GJK.c:663-706) decides -- also straight after the written-out two-point default step"""
    rng = np.random.default_rng(0)
    backups = 0
    for t in range(600):
        n = int(rng.choice([5, 8, 20, 60]))
        scale = 10.0 ** rng.integers(6, 9)
        off = rng.normal(size=3) * scale
        pts = rng.normal(size=(n, 3)) * rng.uniform(1e-3, 1.0) + off
        if t % 4 == 1:
            pts[:, 2] = pts[0, 2]
        if t % 4 == 2:
            pts = np.round(pts, 2)
        vrel = off + rng.normal(size=3) * rng.uniform(1e-3, 2.0)
        want = oracle.gjk(vrel, pts)
        if want[3][2] == 0 and t % 5:
            continue
        backups += int(want[3][2])
        for cap, lanes in ((256, 1), (256, 0)):
            g = host.gjk(vrel, pts, cap=cap, lanes=lanes)
            assert _same_gjk(g, want, n), (t, n, cap, lanes, g, want)
    assert backups >= 10, backups  # st[2]: runs of the backup procedure
