"""Parity on the configuration the headline number is quoted on (BASELINE config 5: 65,536 agents, K = 64, H = 20,
P = 50), against the oracle, bit for bit: a random sample of agents with ALL their 64 slots (status, reachable count,
half-plane, new velocity) and EVERY inside pair of the step (all hulls: facet count, plane).  The pair path compared is
MAIN:1381-1403, the LP MAIN:1101-1112.  Also the race the chunked neighbour upload could run into (host-buffer steps
of >= 2^20 pairs launch the pair kernel per agent block on concurrent streams)."""
import numpy as np
import pytest

from conftest import make_ctx
from lqr_obstacles_b200 import swarm

pytestmark = pytest.mark.gpu


def _oracle_rows(oracle, NF, T, sph, x, vpref, nbr, agents, nthreads=8):
    """oracle step of the listed agents only: their states first, every state they reference after (ids remapped)"""
    agents = np.asarray(agents)
    rows = nbr[agents]
    others = np.unique(rows[rows >= 0])
    ids = np.concatenate([agents, others])
    remap = {int(g): len(agents) + t for t, g in enumerate(others)}
    for t, g in enumerate(agents):
        remap[int(g)] = t                      # an agent of the sample that is also somebody's neighbour
    rr = np.array([[remap[int(j)] if j >= 0 else -1 for j in r] for r in rows], dtype=np.int32)
    xs = np.ascontiguousarray(x[ids])
    vs = np.zeros((len(ids), 3))
    vs[:len(agents)] = vpref[agents]
    full = np.full((len(ids), nbr.shape[1]), -1, dtype=np.int32)
    full[:len(agents)] = rr
    newV, planes, status, nplanes = oracle.step(NF, T, sph, xs, vs, full, nthreads=nthreads)
    n = len(agents)
    return newV[:n], planes[:n], status[:n], nplanes[:n]


def test_swarm65536_sample_and_every_hull_against_the_oracle(golden, oracle, product_lib):
    x, vpref = swarm.lattice_swarm((64, 32, 32))
    K = 64
    nbr = swarm.knn(x[:, :3], K)
    ctx = make_ctx(golden, n_agents=65536, max_neighbors=K, horizon=20)
    newV = ctx.step(x, vpref, nbr)
    assert ctx.last_step_nbr_streamed()   # host table of >= 2^21 pairs: its rows arrived while the pair kernel ran
    res = ctx.pair_results()
    cnt = ctx.counters()
    NF, T, sph = ctx.tables()
    st = res["status"] & 0xFF
    assert cnt["hull_overflow"] == 0 and cnt["degenerate"] == 0
    # (1) 320 random agents x all 64 slots + their LP
    rng = np.random.default_rng(65536)
    agents = np.sort(rng.choice(65536, 320, replace=False))
    oV, opl, ost, onp = _oracle_rows(oracle, NF, T, sph, x, vpref, nbr, agents)
    assert np.array_equal(st[agents], ost & 0xFF)
    has = (ost & 0xFF) != 0
    assert np.array_equal(res["planes"][agents][has], opl[has])
    assert np.array_equal(newV[agents], oV)
    # (2) the agents that own an inside pair (their LP includes a hull plane), whole rows again
    inside = np.argwhere(st == 2)
    assert len(inside) == cnt["inside"] and len(inside) >= 100
    owners = np.unique(inside[:, 0])
    oV2, opl2, ost2, _ = _oracle_rows(oracle, NF, T, sph, x, vpref, nbr, owners)
    assert np.array_equal(st[owners], ost2 & 0xFF)
    has2 = (ost2 & 0xFF) != 0
    assert np.array_equal(res["planes"][owners][has2], opl2[has2])
    assert np.array_equal(newV[owners], oV2)
    # (3) every inside pair on its own: reachable count, facet count, escape distance
    for i, kk in inside:
        o = oracle.pair(NF, T, sph, x[i], x[nbr[i, kk]])
        assert o["status"] == 2 and o["n_reach"] == res["nreach"][i, kk] and o["n_facets"] == res["nfacets"][i, kk], (i, kk)
        assert np.array_equal(o["plane"], res["planes"][i, kk])
        assert abs(o["distance"] - res["dist"][i, kk]) <= 1e-12 * max(1.0, abs(o["distance"]))
    ctx.close()


def test_every_launch_shape_with_overflowing_clouds(golden, product_lib, monkeypatch):
    """Three ways a step of >= 2^20 pairs is launched must give the same bits on a TIGHT swarm -- reachable clouds larger than
    the 256-point shared part, so the per-warp overflow areas in HBM are live, and thousands of hulls:
    (a) default: polling hull launch on its own stream and SMs as a concurrent consumer of the queue, pair kernel, wide hull launch
        -- with a host table its rows are uploaded NEXT TO the pair kernel, which walks the first eighth of the agents first and
        waits for the rows of a batch (lqro_last_step_nbr_streamed); with the resident table there is nothing to wait for;
    (b) LQRO_FLAG_NO_OVERLAP with a host table: the table goes up in agent blocks, one pair launch per block on concurrent streams;
    (c) LQRO_FLAG_NO_OVERLAP with the resident table: one pair launch, hulls afterwards."""
    from lqr_obstacles_b200.api import FLAG_NO_OVERLAP

    x, vpref = swarm.lattice_swarm((32, 32, 16))
    x[:, :3] *= 0.45
    K = 64
    nbr = swarm.knn(x[:, :3], K)
    N = len(x)
    assert N * K >= 1 << 20
    ctx = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20, flags=FLAG_NO_OVERLAP)
    v_chunked = ctx.step(x, vpref, nbr).copy()          # (b)
    r_chunked = ctx.pair_results()
    assert ctx.counters()["launches"] > 4
    assert (r_chunked["nreach"] > 256).sum() > 1000     # the overflow areas were in use
    v_plain = ctx.step(x, vpref, None).copy()           # (c)
    r_plain = ctx.pair_results()
    assert ctx.counters()["launches"] == 4
    ctx.close()
    monkeypatch.setenv("LQRO_NBR_STREAM_MIN_PAIRS", str(1 << 17))   # default: 2^21 pairs
    ctx = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20)
    monkeypatch.delenv("LQRO_NBR_STREAM_MIN_PAIRS")
    v_two = ctx.step(x, vpref, nbr).copy()              # (a), host table: its rows arrive while the pair kernel runs
    r_two = ctx.pair_results()
    cnt = ctx.counters()
    assert cnt["launches"] == 4 and cnt["inside"] > 500
    assert ctx.last_step_nbr_streamed()
    v_res = ctx.step(x, vpref, None).copy()             # (a), resident table: one segment, nothing to wait for
    r_res = ctx.pair_results()
    assert not ctx.last_step_nbr_streamed()
    nbr2 = nbr.copy()
    nbr2[:, 1:] = nbr[:, :0:-1]                         # another table (slots 1.. reversed): the streamed rows are the ones used
    v_rev = ctx.step(x, vpref, nbr2).copy()
    r_rev = ctx.pair_results()
    assert ctx.last_step_nbr_streamed()
    assert np.array_equal(r_rev["status"][:, 1:], r_plain["status"][:, :0:-1]) and np.array_equal(r_rev["nreach"][:, 1:], r_plain["nreach"][:, :0:-1])
    has_rev = (r_rev["status"] & 0xFF) != 0
    assert np.array_equal(r_rev["planes"][:, 1:][has_rev[:, 1:]], r_plain["planes"][:, :0:-1][has_rev[:, 1:]])
    v_two2 = ctx.step(x, vpref, nbr).copy()             # and the first one again
    assert np.array_equal(v_two2, v_plain)
    ctx.close()
    has = (r_plain["status"] & 0xFF) != 0
    for v, r in ((v_chunked, r_chunked), (v_two, r_two), (v_res, r_res)):
        assert np.array_equal(v, v_plain)
        for k in ("status", "nreach", "nfacets"):
            assert np.array_equal(r[k], r_plain[k]), k
        assert np.array_equal(r["planes"][has], r_plain["planes"][has])
