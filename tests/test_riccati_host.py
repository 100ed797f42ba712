"""CPU check of the batched Riccati kernel's LOGIC: the kernel source csrc/lqro_riccati.cuh (riccati_warp) compiled as plain C++
on the 32-thread lockstep emulation of the warp primitives (tests/hull_host/warp_emul.h) against the oracle's restatement of
controlMatrices (MAIN:512-543) and the gains recorded from the reference.  The GPU test (tests/test_gpu_parity.py) runs the same
source on the device.  Tolerance 1e-9 relative (north_star), as on the device."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "riccati_host", "riccati_host.cpp")
SO = os.path.join(HERE, "riccati_host", "libriccatihost.so")
DEPS = [SRC, os.path.join(HERE, "hull_host", "warp_emul.h"), os.path.join(HERE, "..", "lqr_obstacles_b200", "csrc", "lqro_riccati.cuh")]


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in DEPS):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-o", SO, SRC])
    lib = ctypes.CDLL(SO)
    lib.riccatihost_solve.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int] + [ctypes.c_void_p] * 4

    def solve(A, B, qv=100.0, qp=0.1, r=5.0, iters=300):
        A = np.ascontiguousarray(A, dtype=np.float64); B = np.ascontiguousarray(B, dtype=np.float64)
        L, E, Lh, Eh = np.zeros((4, 16)), np.zeros((4, 3)), np.zeros((3, 16)), np.zeros((3, 3))
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        lib.riccatihost_solve(P(A), P(B), qv, qp, r, iters, P(L), P(E), P(Lh), P(Eh))
        return L, E, Lh, Eh

    return solve


def _close(a, b, tol=1e-9):
    return np.abs(a - b).max() <= tol * np.abs(b).max()


def test_reference_gains(host, golden):
    L, E, Lh, Eh = host(golden["dyn_A"], golden["dyn_B"])
    assert _close(L, golden["dyn_L"]) and _close(E, golden["dyn_E"]) and _close(Lh, golden["dyn_Lh"]) and _close(Eh, golden["dyn_Eh"])


def test_other_weights_and_iteration_counts(host, oracle, golden):
    """the iteration count is part of the contract (the position recursion has not converged after 300 iterations)"""
    for qv, qp, r, it in ((100.0, 0.1, 5.0, 7), (30.0, 1.0, 2.0, 120), (250.0, 0.02, 9.0, 400)):
        got = host(golden["dyn_A"], golden["dyn_B"], qv, qp, r, it)
        want = oracle.riccati(golden["dyn_A"], golden["dyn_B"], qv, qp, r, it)
        for g, w_ in zip(got, want):
            assert _close(g, np.asarray(w_).reshape(g.shape)), (qv, qp, r, it)


def test_perturbed_dynamics(host, oracle, golden):
    rng = np.random.default_rng(4)
    for _ in range(3):
        A = golden["dyn_A"] * (1 + 0.05 * rng.normal(size=(16, 16)))
        B = golden["dyn_B"] * (1 + 0.05 * rng.normal(size=(16, 4)))
        got = host(A, B)
        want = oracle.riccati(A, B)
        for g, w_ in zip(got, want):
            assert _close(g, np.asarray(w_).reshape(g.shape))
