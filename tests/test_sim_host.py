"""Simulator-form agent update (MAIN:1408-1417) -- CPU-side checks.

The kernel source csrc/lqro_sim.cuh also compiles as plain C++ (tests/sim_host); that build must agree BIT FOR BIT with
the compiled reference (all arithmetic in the reference's order, the same libm), which pins the kernel's logic without a
GPU.  The GPU parity proper is tests/test_gpu_sim.py (1e-9: CUDA's tan/asin/atan2/log/hypot differ from glibc's in the
last bit).  Also here: lqro_msvc_normals against the reference's normal(), the default model against the reference's
globals, the golden fixture against the reference."""
import ctypes
import os

import numpy as np
import pytest

from _libs import P, RefSim, SimHost

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("x", "Rot", "P", "xTrue", "RotTrue", "vGoal")


@pytest.fixture(scope="module")
def sim_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "sim_update.npz"))


@pytest.fixture(scope="module")
def simhost():
    return SimHost()


def _before(g, t):
    return {k: g["s%d_before_%s" % (t, k)] for k in ("x", "Rot", "P", "xTrue", "RotTrue", "pGoal")}


def test_host_build_of_the_kernel_reproduces_the_golden_updates_bit_for_bit(sim_golden, simhost):
    g = sim_golden
    for t in g["record_steps"]:
        o = simhost.update(g["model"], _before(g, t), g["s%d_newV" % t], g["s%d_normals" % t])
        for k in KEYS:
            assert np.array_equal(o[k], g["s%d_after_%s" % (t, k)]), (int(t), k)


def test_general_path_of_the_flow_also_reproduces_the_golden_updates(sim_golden, simhost):
    """sim_flow with the shared Pade approximant switched off (two separate exponentials, exp(dt F) parked outside the
    workspace, dt/2 F recovered from the scaled copy): the path inputs with |dt F|_1 < 0.95 would take"""
    g = sim_golden
    simhost.lib.simhost_force_general_exp(1)
    try:
        for t in (0, 90):
            o = simhost.update(g["model"], _before(g, t), g["s%d_newV" % t], g["s%d_normals" % t])
            for k in KEYS:
                assert np.array_equal(o[k], g["s%d_after_%s" % (t, k)]), (int(t), k)
    finally:
        simhost.lib.simhost_force_general_exp(0)


def test_diagonal_noise_covariance_shortcut_changes_no_bit(sim_golden, simhost):
    """X*M for the diagonal M = 1e-9 I as one product per element (sim_flow) vs the full 16-term sums: same bits, and both
    equal the reference's update (the default run of the other tests uses the shortcut)"""
    g = sim_golden
    simhost.lib.simhost_use_m_diag(0)
    try:
        for t in (1, 90, 209):
            o = simhost.update(g["model"], _before(g, t), g["s%d_newV" % t], g["s%d_normals" % t])
            for k in KEYS:
                assert np.array_equal(o[k], g["s%d_after_%s" % (t, k)]), (int(t), k)
                assert np.array_equal(np.signbit(o[k]), np.signbit(g["s%d_after_%s" % (t, k)])), (int(t), k)  # zeros keep their sign too
    finally:
        simhost.lib.simhost_use_m_diag(1)
    o = simhost.update(g["model"], _before(g, 90), g["s90_newV"], g["s90_normals"])
    for k in KEYS:
        assert np.array_equal(np.signbit(o[k]), np.signbit(g["s90_after_%s" % k])), k


def test_stage_pieces_match_the_reference_bit_for_bit(reference, simhost):
    rng = np.random.default_rng(7)
    lib, sh = reference.lib, simhost.lib
    for it in range(40):
        q = rng.normal(size=(16, 16)) * rng.choice([0.0, 0.01, 0.1, 1, 5])
        a, b = np.zeros((16, 16)), np.zeros((16, 16))
        lib.ref_exp16(P(q), P(a)); sh.simhost_exp16(P(q), P(b))
        assert np.array_equal(a, b)
        q3 = rng.normal(size=(3, 3)) * rng.choice([0.0, 1e-3, 1, 5])
        a3, b3 = np.zeros((3, 3)), np.zeros((3, 3))
        lib.ref_exp3(P(q3), P(a3)); sh.simhost_exp3(P(q3), P(b3))
        assert np.array_equal(a3, b3)
        m = rng.normal(size=(16, 16))
        m = m @ m.T * 1e-10
        if it % 4 == 0:
            m = np.diag(np.diag(m))  # already diagonal: no rotation at all
        V1, D1, V2, D2 = (np.zeros((16, 16)) for _ in range(4))
        lib.ref_jacobi16(P(m), P(V1), P(D1)); sh.simhost_jacobi16(P(m), P(V2), P(D2))
        assert np.array_equal(V1, V2) and np.array_equal(np.diag(D1), np.diag(D2))


def test_shared_pade_for_exp_and_exp_of_half_is_exact(reference, simhost):
    """exp(q) and exp(q/2) from ONE Pade approximant + solve (csrc/lqro_sim.cuh: sim_exp_with_half) equal the reference's two
    separate exp calls bit for bit, on the shared path and on the fallback (s(q) = 0), incl. norms on the ceil boundary"""
    rng = np.random.default_rng(4)
    taken = [0, 0]
    for it in range(200):
        q = rng.normal(size=(16, 16)) * rng.choice([0.001, 0.02, 0.05, 0.06, 0.1, 0.3, 1, 3, 10])
        if it % 7 == 0:
            q = q / np.abs(q).sum(0).max() * 0.9504178996162932 * rng.choice([1, 2, 4])
        q = np.ascontiguousarray(q)
        f1, h1, f2, h2 = (np.zeros((16, 16)) for _ in range(4))
        reference.lib.ref_exp16(P(q), P(f1))
        reference.lib.ref_exp16(P(np.ascontiguousarray(q * 0.5)), P(h1))
        taken[simhost.lib.simhost_exp_with_half(P(q), P(f2), P(h2))] += 1
        assert np.array_equal(f1, f2) and np.array_equal(h1, h2), it
    assert taken[0] > 10 and taken[1] > 10


def test_model_f_matches_the_reference(reference, simhost, sim_golden):
    rng = np.random.default_rng(11)
    md = np.ascontiguousarray(sim_golden["model"])
    for it in range(50):
        x = rng.normal(size=16)
        if it % 3 == 0:
            x[6:9] = 0.0  # the state the simulator actually evaluates f at (rotation error folded into R)
        x[12:] = 1.2 + 0.1 * rng.normal(size=4)
        Rm = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        u = 1.2 + 0.1 * rng.normal(size=4)
        a, b = np.zeros(16), np.zeros(16)
        reference.lib.ref_f(P(x), P(np.ascontiguousarray(Rm)), P(u), P(a))
        simhost.lib.simhost_f(P(md), P(x), P(np.ascontiguousarray(Rm)), P(u), P(b))
        assert np.array_equal(a, b)


def test_closed_loop_against_the_reference_40_steps(reference, simhost, product_lib):
    """reference avoidance + (host build of the kernel) update == reference avoidance + reference update, bit for bit,
    with the noise drawn by lqro_msvc_normals from the same stream position"""
    from lqr_obstacles_b200.api import msvc_normals
    from golden.make_golden import hexagon

    starts, goals = hexagon()
    sim = RefSim(reference, starts, goals, seed=1)
    md = sim.model()
    n = sim.n
    for step in range(40):
        newV = sim.avoid(H=30)[0]
        st = sim.full_state()
        normals, seed_after = msvc_normals(sim.seed(), 22 * n)
        sim.update()
        assert sim.seed() == seed_after
        ref_after = sim.full_state()
        o = simhost.update(md, st, newV, normals)
        for k in KEYS:
            assert np.array_equal(o[k], ref_after[k]), (step, k)
    sim.close()


def _operating_points(rng, u0, n):
    pts = [(np.concatenate([np.zeros(12), np.full(4, u0)]), np.eye(3).ravel().copy(), np.full(4, u0))]  # controlMatrices' hover point
    for _ in range(n - 1):
        u = u0 + 0.05 * rng.normal(size=4)
        x = rng.normal(size=16) * 0.3
        x[12:] = u0 + 0.05 * rng.normal(size=4)
        R = np.linalg.qr(np.eye(3) + 0.1 * rng.normal(size=(3, 3)))[0].ravel().copy()
        pts.append((x, R, u))
    return pts


def test_linearize_discretize_matches_the_reference_and_the_golden_dynamics(reference, simhost, sim_golden, golden):
    md = np.ascontiguousarray(sim_golden["model"])
    for it, (x, R, u) in enumerate(_operating_points(np.random.default_rng(2), reference.nominal_input(), 8)):
        A1, B1, c1, A2, B2, c2 = (np.zeros(s) for s in [(16, 16), (16, 4), 16] * 2)
        reference.lib.ref_linearize_discretize(P(x), P(R), P(u), P(A1), P(B1), P(c1))
        simhost.lib.simhost_linearize(P(md), P(x), P(R), P(u), P(A2), P(B2), P(c2))
        assert np.array_equal(A1, A2) and np.array_equal(B1, B2) and np.array_equal(c1, c2), it
        if it == 0:  # = the A, B, c controlMatrices produced for the golden vectors
            assert np.array_equal(A2, golden["dyn_A"]) and np.array_equal(B2, golden["dyn_B"]) and np.array_equal(c2, golden["dyn_c"].ravel())


def test_initial_true_states_match_the_reference_setup(simhost, sim_golden, product_lib):
    """setupQuadrotors' sampleGaussian + rotation fold and the first findVGoal (MAIN:126-129, 1355) from rand() state 1"""
    from lqr_obstacles_b200.api import msvc_normals

    g = sim_golden
    n = len(g["init_x"])
    nrm, seed = msvc_normals(1, 16 * n)
    assert seed == int(g["seeds"][0])  # the stream position the reference reaches after setup
    xT, RT, vG = np.zeros((n, 16)), np.zeros((n, 9)), np.zeros((n, 3))
    c = np.ascontiguousarray
    simhost.lib.simhost_init(n, P(c(g["model"])), P(c(g["init_x"])), P(c(g["init_Rot"])), P(c(g["init_P"])), P(c(g["init_pGoal"])), P(nrm),
                             P(xT), P(RT), P(vG))
    assert np.array_equal(xT, g["init_xTrue"]) and np.array_equal(RT, g["init_RotTrue"]) and np.array_equal(vG, g["init_vGoal"])


def test_msvc_normals_are_the_reference_normals(reference, product_lib):
    from lqr_obstacles_b200.api import msvc_normals

    reference.lib.ref_srand(12345)
    a = np.zeros(500)
    reference.lib.ref_normals(500, P(a))
    b, seed = msvc_normals(12345, 500)
    assert np.array_equal(a, b)
    assert seed == reference.lib.ref_get_seed()


def test_default_model_is_the_reference_globals(sim_golden, product_lib):
    from lqr_obstacles_b200.api import sim_default_model

    g = sim_golden["model"]
    m = sim_default_model().as_array()
    gains = slice(7 + 18 + 4, 7 + 18 + 4 + 64 + 12 + 4 + 48 + 9)
    assert np.array_equal(m[:29], g[:29])          # dt .. u_goal
    assert np.array_equal(m[gains.stop:], g[gains.stop:])  # M, N
    assert not m[gains].any()


def test_golden_fixture_is_what_the_reference_computes(reference, sim_golden):
    """regenerates the first recorded steps from oracle/_ref and compares with the committed fixture"""
    from golden.make_golden import hexagon

    g = sim_golden
    starts, goals = hexagon()
    sim = RefSim(reference, starts, goals, seed=1)
    assert np.array_equal(sim.model(), g["model"])
    for t in range(6):
        sim.avoid(H=30)
        assert sim.seed() == int(g["seeds"][t])
        sim.update()
        st = sim.full_state()
        assert np.array_equal(st["xTrue"], g["traj_xTrue"][t])
        if t in (0, 1, 5):
            for k in KEYS:
                assert np.array_equal(st[k], g["s%d_after_%s" % (t, k)])
    sim.close()
