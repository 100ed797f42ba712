"""Parity tests proper: the CUDA path (through the C ABI of liblqro.so) against the CPU oracle and the
golden vectors produced by the reference's own code.  Bars (north_star): bit-exact for discrete results
(reachable sets, statuses, hull facet lists) -- and, because the kernels replay the reference's
arithmetic without FMA, bit-exact for FP64/FP32 values too, except the Riccati gains (1e-9 relative).
"""
import numpy as np
import pytest

from conftest import make_ctx, state_pair_random

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx30(golden, product_lib):
    c = make_ctx(golden, n_agents=6, max_neighbors=5, horizon=30)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx20(golden, product_lib):
    c = make_ctx(golden, n_agents=6, max_neighbors=5, horizon=20)
    yield c
    c.close()


def test_tables_bit_exact(ctx30, oracle, golden):
    """findFG x H and Transform_k (MAIN:653-662, 684) built on the device == reference's F_k, G_k"""
    NF, T, sph = ctx30.tables()
    assert np.array_equal(NF, -golden["F_k"][:, :3, :])
    oNF, oT, _ = oracle.horizon(golden["dyn_A"], golden["dyn_B"], golden["dyn_L"], golden["dyn_E"], 30)
    assert np.array_equal(T, oT)
    assert np.array_equal(sph, golden["spheres"])


def test_riccati_gains(ctx30, golden):
    """controlMatrices (MAIN:512-543) on the device: 1e-9 relative (stated tolerance), batch of 3"""
    A = np.stack([golden["dyn_A"]] * 3)
    B = np.stack([golden["dyn_B"]] * 3)
    L, E, Lh, Eh = ctx30.solve_gains(A, B, 100.0, 0.1, 5.0, 300)
    for b in range(3):
        for got, ref in ((L[b], golden["dyn_L"]), (E[b], golden["dyn_E"]), (Lh[b], golden["dyn_Lh"]), (Eh[b], golden["dyn_Eh"])):
            assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()


def test_riccati_other_weights(ctx30, oracle, golden):
    for qv, qp, r in ((30.0, 0.1, 10.0), (100.0, 1.0, 5.0)):  # tuning variants of the older simulators
        L, E, Lh, Eh = ctx30.solve_gains(golden["dyn_A"], golden["dyn_B"], qv, qp, r, 400)
        oL, oE, oLh, oEh = oracle.riccati(golden["dyn_A"], golden["dyn_B"], qv, qp, r, 400)
        for got, ref in ((L[0], oL), (E[0], oE), (Lh[0], oLh), (Eh[0], oEh)):
            assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()


def test_stage_candidates_reachable_gjk(ctx30, ctx20, oracle, golden):
    """createObstacle, findReachableObstacle, run_gjk per stage on the golden pairs"""
    sph = oracle.spheres()
    for c in range(len(golden["pair_H"])):
        H = int(golden["pair_H"][c])
        ctx = ctx30 if H == 30 else ctx20
        NF, T, _ = ctx.tables()
        xi, xj = golden["pair_xi"][c], golden["pair_xj"][c]
        q = ctx.createObstacle(xi, xj)
        assert np.array_equal(q, oracle.candidates(NF, T, sph, xi, xj))
        idx = ctx.findReachableObstacle(q, xi, xj)
        nre = int(golden["pair_nreach"][c])
        assert np.array_equal(idx, golden["pair_reach_idx"][c][:nre])
        if nre > 4:
            vrel = xi[3:6] - xj[3:6]
            sq, w2, it = ctx.run_gjk(vrel, q[idx])
            d2, ow1, ow2, st = oracle.gjk(vrel, q[idx])
            assert sq == d2 and np.array_equal(w2, ow2) and it == st[0] == golden["pair_gjk"][c][0]
            assert np.array_equal(ctx.last_gjk_witness_vrel, ow1)


def test_gjk_random_clouds(ctx30, oracle):
    rng = np.random.default_rng(5)
    for t in range(60):
        n = int(rng.integers(1, 700))
        pts = rng.normal(size=(n, 3)) * rng.uniform(0.5, 20) + rng.normal(size=3) * 5
        vrel = rng.normal(size=3) * rng.uniform(0.1, 30) if t % 4 else pts.mean(axis=0)
        sq, w2, it = ctx30.run_gjk(vrel, pts)
        d2, ow1, ow2, st = oracle.gjk(vrel, pts)
        assert sq == d2 and np.array_equal(w2, ow2) and it == st[0]


def test_round6(ctx30, oracle):
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.normal(size=4000) * 10.0 ** rng.integers(-12, 9, size=4000),
                        [0.0, 1.0, -2.5, 1.015625, 123456.5, 999999.5, 1e-5, 99999.95, 0.0999999949, 1234565.0, -31.41585]])
    got = ctx30.round6(v)
    want = np.array([float("%g" % x) for x in v])
    assert np.array_equal(got, want)


def test_hull_qhull_fixture(ctx30, golden):
    """reference's own qhull round trip: same facets, same order, same vertex order"""
    planes, verts, deg = ctx30.convex_hull(golden["qh_points"])
    assert deg == 0
    assert np.array_equal(verts, golden["qh_facets"])
    assert np.abs(planes - golden["qh_planes"]).max() < 1e-13


def test_hull_random_clouds_vs_oracle(ctx30, oracle):
    rng = np.random.default_rng(0)
    for t in range(40):
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
        else:
            k = rng.integers(0, 12, size=n)
            d = rng.normal(size=(n, 3))
            d /= np.linalg.norm(d, axis=1)[:, None]
            p = d * (np.array([0.5, 0.5, 1.5]) * (1 + k[:, None] * 0.7)) + k[:, None] * np.array([1.5, 0.3, 0.1])
        p = oracle.round6(p)
        oplanes, overts, order, info = oracle.hull(p)
        planes, verts, deg = ctx30.convex_hull(p)
        assert np.array_equal(verts, overts), (t, n, kind)   # facet list, order and vertex order
        assert np.array_equal(planes, oplanes)                # same formula, no FMA: identical bits
        assert deg == (1 if info["degenerate"] else 0)


def test_hull_escape_vs_oracle(ctx30, oracle):
    """convexHull (MAIN:809-838): 6-digit cloud, nearest facet, distance from the un-rounded first vertex"""
    rng = np.random.default_rng(2)
    for t in range(25):
        n = int(rng.integers(50, 1200))
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1)[:, None]
        pts = d * [6, 7, 15] * rng.uniform(0.6, 1.0, size=(n, 1)) + rng.normal(size=3)
        vrel = rng.normal(size=3)
        nf, od, on, obest, odeg = oracle.hull_escape(pts, vrel, fix_facet0=1)
        dist, nrm, best, nfac = ctx30.convexHull(pts, vrel)
        assert (nfac, best) == (nf, obest)
        assert abs(dist - od) <= 1e-12 * max(1.0, abs(od))   # oracle applies the '%6.16g' text round trip to the normal
        assert np.abs(nrm - on).max() <= 1e-15


def test_lp_golden_bit_exact(ctx30, golden):
    """calculateNewV (MAIN:1101-1112), FP32, batch of 200 ragged cases incl. linearProgram4 fallbacks"""
    out = ctx30.calculateNewV(golden["lp_planes"], golden["lp_n"], golden["lp_goal"])
    assert np.array_equal(out, golden["lp_newv"])


def test_lp_random_vs_oracle(ctx30, oracle):
    rng = np.random.default_rng(7)
    B, M = 512, 64
    planes = np.zeros((B, M, 6), dtype=np.float32)
    counts = np.zeros(B, dtype=np.int32)
    goal = np.zeros((B, 3))
    for t in range(B):
        n = int(rng.integers(0, M + 1))
        normals = rng.normal(size=(M, 3))
        normals /= np.linalg.norm(normals, axis=1)[:, None]
        pts = rng.normal(size=(M, 3)) * rng.uniform(0.1, 5)
        if t % 3 == 0:
            pts = -normals * rng.uniform(0, 3, size=(M, 1))
        if t % 5 == 0 and n > 3:
            normals[2] = -normals[0]
            pts[2] = pts[0] + normals[0] * rng.uniform(0.1, 2)
        planes[t] = np.concatenate([pts, normals], axis=1)
        counts[t] = n
        goal[t] = rng.normal(size=3) * rng.uniform(0.1, 150)
    out = ctx30.calculateNewV(planes, counts, goal)
    for t in range(B):
        assert np.array_equal(out[t], oracle.new_velocity(planes[t, : counts[t]], goal[t])), t


def test_agent_update(ctx30, oracle):
    rng = np.random.default_rng(1)
    pos = rng.normal(size=(1000, 3)).astype(np.float32)
    vel = rng.normal(size=(1000, 3)).astype(np.float32)
    nv = rng.normal(size=(1000, 3)).astype(np.float32)
    p2, v2 = ctx30.agent_update(pos, vel, nv, np.float32(1 / 30))
    op, ov = pos.copy(), vel.copy()
    from _libs import P
    oracle.lib.lqro_o_agent_update(1000, P(op), P(ov), P(nv), np.float32(1 / 30))
    assert np.array_equal(p2, op) and np.array_equal(v2, ov)


def _step_vs_oracle(ctx, oracle, x, vpref, nbr, H):
    NF, T, sph = ctx.tables()
    newV = ctx.step(x, vpref, nbr)
    res = ctx.pair_results()
    oV, opl, ost, onp = oracle.step(NF, T, sph, x, vpref, nbr, nthreads=8)
    st = res["status"] & 0xFF
    assert np.array_equal(st, ost)
    has = ost != 0
    assert np.array_equal(res["planes"][has], opl[has])
    assert np.array_equal(newV, oV)
    return res, ost


def test_six_quad_scenario_steps(golden, oracle, product_lib):
    """Canonical 6-quad scenario: device step == reference loop (status, planes, newV) on recorded states.
    Pair-steps hit by the reference's stale-normal defect (facet 0 nearest) are excluded, and counted."""
    from lqr_obstacles_b200.swarm import all_pairs_neighbors

    ctx = make_ctx(golden, n_agents=6, max_neighbors=5, horizon=30)
    nbr = all_pairs_neighbors(6)
    inside = stale = 0
    for t in range(0, golden["sc6_x_vgoal"].shape[0], 2):
        x = np.ascontiguousarray(golden["sc6_x_vgoal"][t][:, :16])
        vg = np.ascontiguousarray(golden["sc6_x_vgoal"][t][:, 16:])
        newV = ctx.step(x, vg, nbr)
        res = ctx.pair_results()
        st = res["status"] & 0xFF
        assert np.array_equal(st, golden["sc6_status"][t])
        same = np.all(res["planes"] == golden["sc6_planes"][t], axis=2) | (st == 0)
        for i, kk in np.argwhere(~same):
            assert st[i, kk] == 2
            stale += 1
        inside += int((st == 2).sum())
        ok = same.all(axis=1)
        assert np.array_equal(newV[ok], golden["sc6_newv"][t][ok])
    assert inside > 40 and stale <= 3
    ctx.close()


def test_pairs_golden_full_pipeline(golden, product_lib):
    """the 86 golden pairs (outside / inside / none) through lqro_step as 2-agent problems"""
    for H in (20, 30):
        ctx = make_ctx(golden, n_agents=2, max_neighbors=1, horizon=H)
        nbr = np.array([[1], [0]], dtype=np.int32)
        for c in np.where(golden["pair_H"] == H)[0]:
            x = np.ascontiguousarray(np.stack([golden["pair_xi"][c], golden["pair_xj"][c]]))
            ctx.step(x, np.zeros((2, 3)), nbr)
            res = ctx.pair_results()
            st = int(golden["pair_status"][c])
            assert (res["status"][0, 0] & 0xFF) == st
            assert res["nreach"][0, 0] == golden["pair_nreach"][c]
            if st == 1:
                assert res["dist"][0, 0] == golden["pair_distance"][c]
                assert np.array_equal(res["planes"][0, 0], golden["pair_plane"][c])
            if st == 2 and golden["pair_best_facet"][c] != 0:  # facet 0: reference keeps a stale normal
                assert abs(res["dist"][0, 0] - golden["pair_distance"][c]) <= 1e-12 * max(1.0, golden["pair_distance"][c])
                assert np.abs(res["planes"][0, 0] - golden["pair_plane"][c]).max() <= 1e-6  # FP32 storage
        ctx.close()


def test_swarm_step_vs_oracle(golden, oracle, product_lib):
    """synthetic swarm (config 4 recipe at 512 agents, K=32, H=20): whole step == oracle, bit for bit"""
    from lqr_obstacles_b200 import swarm

    x, vpref = swarm.lattice_swarm((8, 8, 8))
    nbr = swarm.knn(x[:, :3], 32)
    ctx = make_ctx(golden, n_agents=512, max_neighbors=32, horizon=20)
    res, ost = _step_vs_oracle(ctx, oracle, x, vpref, nbr, 20)
    assert (ost == 1).sum() > 1000
    # ragged / empty neighbour slots and a dense close-packed variant that exercises the hull branch
    nbr2 = nbr.copy()
    nbr2[::3, 5:] = -1
    nbr2[7] = -1
    _step_vs_oracle(ctx, oracle, x, vpref, nbr2, 20)
    x3 = x.copy()
    x3[:, 0:3] *= 0.22
    x3[:, 3:6] *= 0.3
    res3, ost3 = _step_vs_oracle(ctx, oracle, x3, vpref, swarm.knn(x3[:, :3], 32), 20)
    assert (ost3 == 2).sum() > 50
    c = ctx.counters()
    assert c["inside"] == (ost3 == 2).sum() and c["launches"] == 4 and c["hull_overflow"] == 0
    ctx.close()
    # experimental fused variant (hull warps inside the pair kernel): same bits
    ctx_f = make_ctx(golden, n_agents=512, max_neighbors=32, horizon=20, flags=2)
    _step_vs_oracle(ctx_f, oracle, x3, vpref, swarm.knn(x3[:, :3], 32), 20)
    ctx_f.close()


def test_full_size_properties(golden, oracle, product_lib):
    """BASELINE config 4 at full size (4,096 agents, K=32, H=20): size-independent properties --
    idempotence, slot-permutation equivariance, LP feasibility -- plus a random sample against the oracle."""
    from lqr_obstacles_b200 import swarm

    x, vpref = swarm.lattice_swarm((16, 16, 16))
    nbr = swarm.knn(x[:, :3], 32)
    ctx = make_ctx(golden, n_agents=4096, max_neighbors=32, horizon=20)
    v1 = ctx.step(x, vpref, nbr).copy()
    r1 = ctx.pair_results()
    v2 = ctx.step(x, vpref, nbr).copy()
    r2 = ctx.pair_results()
    assert np.array_equal(v1, v2) and all(np.array_equal(r1[k], r2[k]) for k in r1)       # idempotent / deterministic
    perm = np.random.default_rng(0).permutation(32)
    ctx.step(x, vpref, np.ascontiguousarray(nbr[:, perm]))
    r3 = ctx.pair_results()
    assert np.array_equal(r3["status"], r1["status"][:, perm])
    has3 = (r3["status"] & 0xFF) != 0   # slots without a plane keep whatever an earlier step left there
    assert np.array_equal(r3["planes"][has3], r1["planes"][:, perm][has3])
    # feasibility: newV satisfies every plane of its agent (normal.(v - point) >= -eps) when not in the lp4 fallback
    st = r1["status"] & 0xFF
    pl = r1["planes"]
    viol = np.einsum("nkd,nkd->nk", pl[:, :, 3:], v1[:, None, :].astype(np.float32) - pl[:, :, :3])
    viol = np.where(st != 0, viol, 0.0)
    assert (viol.min(axis=1) > -1e-3).mean() > 0.95
    assert np.all(np.linalg.norm(v1, axis=1) <= 100.0 * (1 + 1e-6))
    # the exact pruning of horizon steps changes nothing: same step with every step evaluated
    ctx_np = make_ctx(golden, n_agents=4096, max_neighbors=32, horizon=20, flags=1)
    v_np = ctx_np.step(x, vpref, nbr)
    r_np = ctx_np.pair_results()
    assert np.array_equal(v_np, v1) and np.array_equal(r_np["status"], r1["status"]) and np.array_equal(r_np["nreach"], r1["nreach"])
    has1 = (r1["status"] & 0xFF) != 0
    assert np.array_equal(r_np["planes"][has1], r1["planes"][has1]) and np.array_equal(r_np["dist"][has1], r1["dist"][has1])
    ctx_np.close()
    # sample against the oracle
    NF, T, sph = ctx.tables()
    rng = np.random.default_rng(1)
    for i in rng.choice(4096, 48, replace=False):
        for kk in rng.choice(32, 4, replace=False):
            o = oracle.pair(NF, T, sph, x[i], x[nbr[i, kk]])
            assert o["status"] == st[i, kk] and o["n_reach"] == r1["nreach"][i, kk]
            if o["status"]:
                assert np.array_equal(o["plane"], pl[i, kk])
    ctx.close()
