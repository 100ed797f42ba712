"""Trajectory equivalence (BASELINE configs 1-3): the reference simulator's own update code (LQR feedback,
nonlinear propagation with the emulated MSVC noise stream, Kalman filter -- oracle/_ref) is run twice, once
with the reference's CPU avoidance step and once with the device step (lqro_step) supplying newV.  Needs
the prebuilt oracle/_ref/liblqro_ref.so (it travels to the GPU box; skipped when absent).

Stated tolerance: positions within 1e-6 m over the whole run (they are normally bit-identical; the only
arithmetic the device does not replay is the '%6.16g' text round trip of hull normals, <= 1e-16 relative).
The harness applies the facet-0 normal fix ("stale normal") to the reference loop and counts
the events."""
import numpy as np
import pytest

from conftest import make_ctx

pytestmark = pytest.mark.gpu


def run_pair(reference, golden, starts, goals, steps, H=30):
    from _libs import RefSim
    from lqr_obstacles_b200.swarm import all_pairs_neighbors

    n = len(starts)
    reference.lib.ref_sim_set_fix_facet0(1)
    ev0 = reference.lib.ref_sim_facet0_events()
    # --- reference avoidance
    sim = RefSim(reference, starts, goals, seed=1)
    A, B, L, E, Lh, Eh = sim.dynamics()
    traj_ref, inside_ref = [], 0
    for t in range(steps):
        newV, planes, status, npl, nre = sim.avoid(H, 30.0)
        inside_ref += int((status == 2).sum())
        sim.update()
        traj_ref.append(sim.state()[1].copy())
    sim.close()
    events = reference.lib.ref_sim_facet0_events() - ev0
    # --- device avoidance, same simulator, same noise stream
    ctx = make_ctx(golden, n_agents=n, max_neighbors=n - 1, horizon=H)
    ctx.set_dynamics(A, B, L, E)
    nbr = all_pairs_neighbors(n)
    sim = RefSim(reference, starts, goals, seed=1)
    traj_gpu, inside_gpu, nbit = [], 0, 0
    run_pair.degenerate = 0  # hulls whose merged (non-simplicial) qhull facets the device does not restate (status bit 0x100): must stay 0 here
    for t in range(steps):
        x, xt, vg = sim.state()
        newV = ctx.step(np.ascontiguousarray(x), np.ascontiguousarray(vg), nbr)
        inside_gpu += int(((ctx.pair_results()["status"] & 0xFF) == 2).sum())
        run_pair.degenerate += int(ctx.counters()["degenerate"])
        sim.set_newv(newV)
        sim.update()
        traj_gpu.append(sim.state()[1].copy())
    sim.close()
    ctx.close()
    traj_ref, traj_gpu = np.array(traj_ref), np.array(traj_gpu)
    return traj_ref, traj_gpu, inside_ref, inside_gpu, events


def test_six_quads_1000_steps(reference, golden, product_lib):
    from lqr_obstacles_b200.swarm import hexagon6

    S, G = hexagon6()
    tr, tg, ir, ig, ev = run_pair(reference, golden, S, G, 1000)
    err = np.abs(tr[:, :, :3] - tg[:, :, :3]).max()
    print("6 quads x 1000 steps: max |dp| = %.3e m, inside pair-steps %d / %d, facet-0 events %d, degenerate hulls %d, bit-identical states: %s"
          % (err, ir, ig, ev, run_pair.degenerate, np.array_equal(tr, tg)))
    assert ir == ig and ir > 100
    assert err <= 1e-6
    assert np.abs(tr[-1, :, :3] - G).max() < 0.5  # and the quads did reach their goals


@pytest.mark.parametrize("n,steps", [(4, 1000), (8, 1000), (16, 1000)])  # BASELINE config 2: 1,000-step equivalence
def test_reference_scenarios(reference, golden, product_lib, n, steps):
    from lqr_obstacles_b200.swarm import quad24

    S, G = quad24()
    tr, tg, ir, ig, ev = run_pair(reference, golden, S[:n], G[:n], steps)
    err = np.abs(tr[:, :, :3] - tg[:, :, :3]).max()
    print("%d quads x %d steps: max |dp| = %.3e m, inside pair-steps %d / %d, facet-0 events %d, degenerate hulls %d" % (n, steps, err, ir, ig, ev, run_pair.degenerate))
    assert ir == ig
    assert err <= 1e-6


@pytest.mark.parametrize("n,radius,steps", [(32, 12.0, 60), (64, 24.0, 25)])
def test_large_reference_scenarios(reference, golden, product_lib, n, radius, steps):
    """BASELINE configs 32 and 64 quads (the repo holds no initial conditions for them: circle-antipodal sets of radius
    12 / 24 m); K = N-1 as in the reference, its own H = 30"""
    from lqr_obstacles_b200.swarm import circle_antipodal

    S, G = circle_antipodal(n, radius)
    tr, tg, ir, ig, ev = run_pair(reference, golden, S, G, steps)
    err = np.abs(tr[:, :, :3] - tg[:, :, :3]).max()
    print("%d quads x %d steps: max |dp| = %.3e m, inside pair-steps %d / %d, facet-0 events %d, degenerate hulls %d" % (n, steps, err, ir, ig, ev, run_pair.degenerate))
    assert ir == ig
    assert err <= 1e-6
