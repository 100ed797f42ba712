"""Simulator-form agent update on the device (MAIN:1408-1417) against the reference.

Tolerance: 1e-9 relative to the magnitude of each quantity (north_star's FP64 bar).  Everything in the kernel is the
reference's arithmetic in the reference's order (-fmad=false) except the libm calls (tan, asin, atan2, log, hypot), where
CUDA's results may differ from glibc's in the last bit; observed differences are reported by the tests' assert messages."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("x", "Rot", "P", "xTrue", "RotTrue")
RTOL = 1e-9


@pytest.fixture(scope="module")
def sim_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "sim_update.npz"))


@pytest.fixture(scope="module")
def ctx(sim_golden):
    from lqr_obstacles_b200 import Lqro
    from lqr_obstacles_b200.api import SimModel

    c = Lqro(n_agents=6, max_neighbors=5, horizon=30)
    c.sim_set_model(SimModel.from_array(sim_golden["model"]))
    yield c
    c.close()


def _close(a, b, what):
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err <= RTOL, "%s: relative error %.3e > %.0e" % (what, err, RTOL)
    return err


def _run(ctx, g, t, reps=1):
    st = {k: np.tile(g["s%d_before_%s" % (t, k)], (reps, 1)) for k in KEYS}
    tile = lambda a: np.tile(a, (reps, 1))
    vG, u = ctx.sim_update(st["x"], st["Rot"], st["P"], st["xTrue"], st["RotTrue"], tile(g["s%d_newV" % t]),
                           tile(g["s%d_before_pGoal" % t]), tile(g["s%d_normals" % t]))
    st["vGoal"] = vG
    return st


def test_update_matches_the_golden_reference_updates(ctx, sim_golden):
    g = sim_golden
    worst = 0.0
    for t in g["record_steps"]:
        o = _run(ctx, g, int(t))
        for k in KEYS + ("vGoal",):
            worst = max(worst, _close(o[k], g["s%d_after_%s" % (t, k)], "step %d %s" % (t, k)))
    print("worst relative error over the recorded steps: %.3e" % worst)


def test_many_agents_grid_stride_and_determinism(ctx, sim_golden):
    """2,400 agents (400 copies of the 6 quads at the close-approach step): more warps than one wave, so the grid-stride
    loop and workspace reuse are exercised; every copy must give the same bits"""
    g = sim_golden
    o = _run(ctx, g, 90, reps=400)
    for k in KEYS + ("vGoal",):
        a = o[k].reshape(400, 6, -1)
        assert np.array_equal(a, np.broadcast_to(a[0], a.shape)), k
        _close(a[0], g["s90_after_%s" % k], k)
    assert ctx.sim_last_ms() > 0.0


def test_against_the_host_build_of_the_kernel_source_on_random_states(ctx, sim_golden):
    """no reference needed: the same source compiled as plain C++ (tests/sim_host, bit-identical to the reference where the
    reference is available, tests/test_sim_host.py) on random states, incl. the control input u and odd agent counts
    (7 = one full and one partial group of the 4-agents-per-warp noise kernel; 1)"""
    from _libs import SimHost

    sh = SimHost()
    g = sim_golden
    rng = np.random.default_rng(5)
    for n in (7, 1):
        x = rng.normal(size=(n, 16)) * 0.5
        x[:, 6:9] = 0.0
        x[:, 12:] = 1.226 + 0.05 * rng.normal(size=(n, 4))
        xT = x + rng.normal(size=(n, 16)) * 1e-4
        xT[:, 6:9] = 0.0
        Rs = np.array([np.linalg.qr(np.eye(3) + 0.1 * rng.normal(size=(3, 3)))[0].ravel() for _ in range(n)])
        Rs *= np.sign(Rs[:, [0]])
        A = rng.normal(size=(n, 16, 16)) * 1e-5
        Pc = (A @ A.transpose(0, 2, 1) + 1e-9 * np.eye(16)).reshape(n, 256)
        st = dict(x=x, Rot=Rs, P=Pc, xTrue=xT, RotTrue=Rs.copy(), pGoal=rng.normal(size=(n, 3)) * 5)
        newV, normals = rng.normal(size=(n, 3)), rng.normal(size=(n, 22))
        want = sh.update(g["model"], st, newV, normals)
        got = {k: st[k].copy() for k in KEYS}
        got["vGoal"], got["u"] = ctx.sim_update(got["x"], got["Rot"], got["P"], got["xTrue"], got["RotTrue"], newV, st["pGoal"], normals)
        for k in KEYS + ("vGoal", "u"):
            _close(got[k], want[k], "n=%d %s" % (n, k))


def test_control_matrices_entirely_on_the_device(ctx, sim_golden, golden):
    """controlMatrices (MAIN:490-544) = lqro_linearize_discretize at the hover point + lqro_solve_gains: A, B, c and the four
    gains against the reference's (golden vectors), 1e-9; plus random operating points against the host build of the kernel"""
    from _libs import P, SimHost

    u0 = float(sim_golden["model"][25])  # u_goal[0] = nominalInput
    x = np.concatenate([np.zeros(12), np.full(4, u0)])
    A, B, c = ctx.linearize_discretize(x, np.eye(3), np.full(4, u0))
    _close(A[0], golden["dyn_A"], "A")
    _close(B[0], golden["dyn_B"], "B")
    assert np.abs(c[0] - golden["dyn_c"].ravel()).max() <= 1e-12
    L, E, Lh, Eh = ctx.solve_gains(A, B)
    for got, k in ((L, "dyn_L"), (E, "dyn_E"), (Lh, "dyn_Lh"), (Eh, "dyn_Eh")):
        _close(got[0], golden[k], k)
    sh = SimHost()
    rng = np.random.default_rng(9)
    n = 37
    xs = rng.normal(size=(n, 16)) * 0.3
    xs[:, 12:] = u0 + 0.05 * rng.normal(size=(n, 4))
    Rs = np.array([np.linalg.qr(np.eye(3) + 0.1 * rng.normal(size=(3, 3)))[0].ravel() for _ in range(n)])
    us = u0 + 0.05 * rng.normal(size=(n, 4))
    A, B, c = ctx.linearize_discretize(xs, Rs, us)
    md = np.ascontiguousarray(sim_golden["model"])
    for i in range(n):
        A2, B2, c2 = np.zeros((16, 16)), np.zeros((16, 4)), np.zeros(16)
        sh.lib.simhost_linearize(P(md), P(np.ascontiguousarray(xs[i])), P(np.ascontiguousarray(Rs[i])), P(np.ascontiguousarray(us[i])), P(A2), P(B2), P(c2))
        _close(A[i], A2, "A[%d]" % i)
        _close(B[i], B2, "B[%d]" % i)
        _close(c[i], c2, "c[%d]" % i)


def test_initial_states_on_the_device(ctx, sim_golden):
    from lqr_obstacles_b200.api import msvc_normals

    g = sim_golden
    nrm, _ = msvc_normals(1, 16 * 6)
    xT, RT, vG = ctx.sim_init(g["init_x"], g["init_Rot"], g["init_P"], g["init_pGoal"], nrm)
    _close(xT, g["init_xTrue"], "xTrue")
    _close(RT, g["init_RotTrue"], "RotTrue")
    _close(vG, g["init_vGoal"], "vGoal")


def test_headless_simulator_reproduces_the_reference_run(sim_golden, product_lib, tmp_path):
    """tools/headless/lqro_headless: the reference's _tmain headless over the C ABI (setup incl. controlMatrices on the device,
    210 steps of the canonical scenario).  Its gains come from the device Riccati solve (1e-14 from the reference's), so
    the trajectory is held to 1e-5 m of the reference's own run; the rand() stream must end where the reference's ends."""
    import json
    import subprocess

    from lqr_obstacles_b200 import build

    exe = build.build_headless()
    out = str(tmp_path / "traj.bin")
    r = subprocess.run([exe, "--scenario", "hexagon6", "--steps", "210", "--out", out, "--quiet"], capture_output=True, text=True, check=True)
    m = json.loads(r.stdout.strip().splitlines()[-1])
    from lqr_obstacles_b200.formats import read_trajectory

    x_true, x_est = read_trajectory(out)
    assert x_true.shape == (210, 6, 16)
    g = sim_golden
    err = np.abs(x_true[:, :, :6] - g["traj_xTrue"][:, :, :6]).max()
    err_est = np.abs(x_est[:, :, :6] - g["traj_x"][:, :, :6]).max()
    print("headless run: max |xTrue - reference| %.3e m, estimates %.3e, %s" % (err, err_est, m))
    assert err <= 1e-5 and err_est <= 1e-5
    assert m["inside_pairs"] > 0 and m["max_distance_to_goal_m"] < 1.0


def test_requires_a_model():
    from lqr_obstacles_b200 import Lqro, LqroError

    c = Lqro(n_agents=6, max_neighbors=5, horizon=30)
    z = lambda *s: np.zeros(s)
    with pytest.raises(LqroError):
        c.sim_update(z(1, 16), z(1, 9), z(1, 256), z(1, 16), z(1, 9), z(1, 3), z(1, 3), z(1, 22))
    c.close()


def test_closed_loop_on_the_device_follows_the_reference_trajectory(sim_golden, golden):
    """Whole time step on the GPU path -- avoidance (lqro_step) + simulator-form update (lqro_sim_update) -- for the
    canonical 6-quad scenario, 210 steps, noise from the reference's stream: the true-state trajectory must stay within
    1e-6 m (absolute) of the reference's own (avoidance decisions are discrete, so last-bit libm differences may be
    amplified; the observed deviation is printed)."""
    from lqr_obstacles_b200 import Lqro
    from lqr_obstacles_b200.api import SimModel, msvc_normals

    g = sim_golden
    n = 6
    c = Lqro(n_agents=n, max_neighbors=n - 1, horizon=30)
    c.set_dynamics(golden["dyn_A"], golden["dyn_B"], golden["dyn_L"], golden["dyn_E"])
    c.sim_set_model(SimModel.from_array(g["model"]))
    st = {k: g["init_" + k].copy() for k in KEYS}
    vpref = g["init_vGoal"].copy()
    pGoal = g["init_pGoal"].copy()
    nbr = np.array([[j for j in range(n) if j != i] for i in range(n)], dtype=np.int32)
    seed = int(g["seeds"][0])
    worst = 0.0
    steps = len(g["traj_xTrue"])
    for t in range(steps):
        newV = c.step(st["x"], vpref, nbr)
        normals, seed = msvc_normals(seed, 22 * n)
        vpref, _ = c.sim_update(st["x"], st["Rot"], st["P"], st["xTrue"], st["RotTrue"], newV, pGoal, normals)
        worst = max(worst, np.abs(st["xTrue"][:, :6] - g["traj_xTrue"][t][:, :6]).max())
    c.close()
    print("max |xTrue - reference| over %d steps: %.3e" % (steps, worst))
    assert worst <= 1e-6


def test_resident_time_steps_equal_the_host_buffer_loop(sim_golden, golden):
    """lqro_timestep_device (avoidance + update, state resident, launches only) over 60 steps of the canonical scenario:
    same bits as the loop of lqro_step + lqro_sim_update with host buffers, and within 1e-6 m of the reference trajectory"""
    import torch
    from lqr_obstacles_b200 import Lqro
    from lqr_obstacles_b200.api import SimModel, msvc_normals

    g = sim_golden
    n, steps = 6, 60
    nbr = np.array([[j for j in range(n) if j != i] for i in range(n)], dtype=np.int32)

    def make():
        c = Lqro(n_agents=n, max_neighbors=n - 1, horizon=30)
        c.set_dynamics(golden["dyn_A"], golden["dyn_B"], golden["dyn_L"], golden["dyn_E"])
        c.sim_set_model(SimModel.from_array(g["model"]))
        return c

    # host-buffer loop
    c = make()
    st = {k: g["init_" + k].copy() for k in KEYS}
    vpref, seed, host_traj = g["init_vGoal"].copy(), int(g["seeds"][0]), []
    for t in range(steps):
        newV = c.step(st["x"], vpref, nbr)
        normals, seed = msvc_normals(seed, 22 * n)
        vpref, _ = c.sim_update(st["x"], st["Rot"], st["P"], st["xTrue"], st["RotTrue"], newV, g["init_pGoal"], normals)
        host_traj.append((st["x"].copy(), st["xTrue"].copy(), vpref.copy()))
    c.close()
    # resident loop: one call per step, only the normals go up
    c = make()
    c.step(g["init_x"].copy(), g["init_vGoal"].copy(), nbr)  # makes the neighbour table (and the state) resident
    c.upload_local_state(np.ascontiguousarray(g["init_x"]))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    d = {k: dev(g["init_" + k]) for k in ("Rot", "P", "xTrue", "RotTrue", "pGoal", "vGoal")}
    dx = torch.empty(0)
    seed = int(g["seeds"][0])
    torch.cuda.synchronize()
    for t in range(steps):
        normals, seed = msvc_normals(seed, 22 * n)
        dn = dev(normals)
        torch.cuda.synchronize()
        c.timestep_device(d["Rot"].data_ptr(), d["P"].data_ptr(), d["xTrue"].data_ptr(), d["RotTrue"].data_ptr(), d["pGoal"].data_ptr(),
                          dn.data_ptr(), d["vGoal"].data_ptr())
        c.synchronize()
        hx, hxt, hv = host_traj[t]
        assert np.array_equal(d["xTrue"].cpu().numpy(), hxt), t
        assert np.array_equal(d["vGoal"].cpu().numpy(), hv), t
        assert np.abs(hxt[:, :6] - g["traj_xTrue"][t][:, :6]).max() <= 1e-6
    c.close()


def test_resident_time_steps_with_device_neighbour_search(sim_golden, golden):
    """lqro_timestep_device with LQRO_FLAG_DEVICE_KNN on a 216-agent lattice swarm, K = 12: neighbour table rebuilt on the
    device from the current positions every step; only the normals go up.  Same bits as the host loop that builds the
    lists with the kd-tree restatement (swarm.knn), steps with lqro_step and updates with lqro_sim_update."""
    import torch
    from lqr_obstacles_b200 import Lqro, swarm
    from lqr_obstacles_b200.api import FLAG_DEVICE_KNN, SimModel

    g = sim_golden
    x0, vpref0 = swarm.lattice_swarm((6, 6, 6))
    x0[:, :3] *= 0.6
    n, K, steps = len(x0), 12, 5
    rng = np.random.default_rng(8)
    normals = rng.normal(size=(steps, n, 22))
    pGoal = -x0[:, :3].copy()
    eye, p0 = np.tile(np.eye(3).ravel(), (n, 1)), np.tile((1e-9 * np.eye(16)).ravel(), (n, 1))

    def make(flags):
        c = Lqro(n_agents=n, max_neighbors=K, horizon=20, flags=flags)
        c.set_dynamics(golden["dyn_A"], golden["dyn_B"], golden["dyn_L"], golden["dyn_E"])
        c.sim_set_model(SimModel.from_array(g["model"]))
        return c

    c = make(0)
    st = dict(x=x0.copy(), Rot=eye.copy(), P=p0.copy(), xTrue=x0.copy(), RotTrue=eye.copy())
    vpref = vpref0.copy()
    for t in range(steps):
        newV = c.step(st["x"], vpref, swarm.knn(st["x"][:, :3], K))
        vpref, _ = c.sim_update(st["x"], st["Rot"], st["P"], st["xTrue"], st["RotTrue"], newV, pGoal, normals[t])
    c.close()
    c = make(FLAG_DEVICE_KNN)
    c.upload_local_state(np.ascontiguousarray(x0))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    d = dict(Rot=dev(eye), P=dev(p0), xTrue=dev(x0), RotTrue=dev(eye), pGoal=dev(pGoal), vGoal=dev(vpref0))
    torch.cuda.synchronize()
    for t in range(steps):
        dn = dev(normals[t])
        torch.cuda.synchronize()
        c.timestep_device(d["Rot"].data_ptr(), d["P"].data_ptr(), d["xTrue"].data_ptr(), d["RotTrue"].data_ptr(), d["pGoal"].data_ptr(),
                          dn.data_ptr(), d["vGoal"].data_ptr())
        c.synchronize()
    assert np.array_equal(d["xTrue"].cpu().numpy(), st["xTrue"])
    assert np.array_equal(d["P"].cpu().numpy(), st["P"])
    assert np.array_equal(d["vGoal"].cpu().numpy(), vpref)
    c.close()


def test_update_against_the_compiled_reference_on_perturbed_states(ctx, reference, sim_golden):
    """states off the canonical trajectory: random attitude errors, velocities and covariances pushed through the
    reference's own update (oracle/_ref) and through the kernel with the same normals"""
    from _libs import RefSim
    from lqr_obstacles_b200.api import msvc_normals
    from golden.make_golden import hexagon

    starts, goals = hexagon()
    sim = RefSim(reference, starts, goals, seed=1)
    rng = np.random.default_rng(3)
    n = sim.n
    for trial in range(6):
        st = sim.full_state()
        x = st["x"] + rng.normal(size=(n, 16)) * 0.05
        x[:, 6:9] = 0.0
        xT = x + rng.normal(size=(n, 16)) * 1e-4
        xT[:, 6:9] = 0.0
        Rs = []
        for i in range(n):
            w = rng.normal(size=3) * 0.2
            W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
            Rs.append((np.eye(3) + W + W @ W / 2).ravel())
        Rs = np.array(Rs)
        A = rng.normal(size=(n, 16, 16)) * 1e-5
        Pc = (A @ A.transpose(0, 2, 1) + 1e-9 * np.eye(16)).reshape(n, 256)
        sim.set_state(x=x, xTrue=xT, Rot=Rs, RotTrue=Rs, Pcov=Pc)
        newV = rng.normal(size=(n, 3))
        sim.set_newv(newV)
        before = sim.full_state()
        normals, seed_after = msvc_normals(sim.seed(), 22 * n)
        sim.update()
        assert sim.seed() == seed_after
        after = sim.full_state()
        o = {k: before[k].copy() for k in KEYS}
        vG, _ = ctx.sim_update(o["x"], o["Rot"], o["P"], o["xTrue"], o["RotTrue"], newV, before["pGoal"], normals)
        o["vGoal"] = vG
        for k in KEYS + ("vGoal",):
            _close(o[k], after[k], "trial %d %s" % (trial, k))
    sim.close()
