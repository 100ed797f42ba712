"""Edge cases of the per-step path on the device: ragged / empty neighbour lists, self slots, agents without any
plane, K = 1, the smallest swarm, a full hull work queue, coplanar clouds, size limits.  Expected values come
from the oracle (bit-exact) or from the reference's own rules (MAIN:1379, 1393, 1101-1112)."""
import numpy as np
import pytest

from conftest import make_ctx
from lqr_obstacles_b200 import LqroError, swarm
from lqr_obstacles_b200.api import PAIR_DROPPED

pytestmark = pytest.mark.gpu


def _tight_swarm(shape=(4, 4, 4), scale=0.5):
    x, vp = swarm.lattice_swarm(shape)
    x[:, :3] *= scale  # closer than the bench lattice: some pairs take the hull branch
    return x, vp


def test_ragged_and_self_slots(golden, oracle, product_lib):
    """-1 slots are empty, a slot naming the agent itself is skipped (MAIN:1379); planes keep slot order"""
    x, vp = _tight_swarm()
    N, K = len(x), 12
    nbr = swarm.knn(x[:, :3], K)
    rng = np.random.default_rng(3)
    nbr[rng.random(nbr.shape) < 0.3] = -1                   # ragged rows
    nbr[::5, 0] = np.arange(N)[::5]                         # self in slot 0
    nbr[7] = -1                                             # an agent with no neighbour at all
    ctx = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20)
    v = ctx.step(x, vp, nbr)
    res = ctx.pair_results()
    NF, T, sph = ctx.tables()
    ov, opl, ost, onp = oracle.step(NF, T, sph, x, vp, nbr, nthreads=4)
    st = res["status"] & 0xFF
    assert np.array_equal(st, ost & 0xFF)
    assert np.array_equal(res["planes"][st != 0], opl[st != 0])
    assert np.array_equal(v, ov)
    assert (st[nbr < 0] == 0).all() and (st[nbr == np.arange(N)[:, None]] == 0).all()
    # no plane at all => linearProgram3 returns the preferred velocity (|vpref| < 100), narrowed to float (V3:113-118)
    assert np.array_equal(v[7], vp[7].astype(np.float32).astype(np.float64))
    ctx.close()


def test_k1_and_two_agents(golden, oracle, product_lib):
    x, vp = _tight_swarm((2, 1, 1), 0.3)
    for K, nbr in ((1, np.array([[1], [0]], dtype=np.int32)), (3, np.array([[1, -1, -1], [-1, 0, -1]], dtype=np.int32))):
        ctx = make_ctx(golden, n_agents=2, max_neighbors=K, horizon=20)
        v = ctx.step(x, vp, nbr)
        NF, T, sph = ctx.tables()
        ov, opl, ost, _ = oracle.step(NF, T, sph, x, vp, nbr)
        assert np.array_equal(v, ov) and np.array_equal(ctx.pair_results()["status"] & 0xFF, ost & 0xFF)
        ctx.close()


def test_hull_queue_overflow_is_flagged(golden, product_lib):
    """max_hull_pairs smaller than the number of inside pairs: lqro_step returns LQRO_E_LIMIT, the extra pairs carry
    LQRO_PAIR_DROPPED and no plane, the counter reports them, the others are unchanged"""
    x, vp = _tight_swarm()
    N, K = len(x), 16
    nbr = swarm.knn(x[:, :3], K)
    full = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20)
    full.step(x, vp, nbr)
    r0, c0 = full.pair_results(), full.counters()
    full.close()
    assert c0["inside"] >= 3 and c0["hull_overflow"] == 0
    small = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20, max_hull_pairs=2)
    with pytest.raises(LqroError, match="hull queue overflow"):   # the host-synchronous step reports it (LQRO_E_LIMIT) ...
        small.step(x, vp, nbr)
    # ... after finishing the step: results and counters are those of the step with the dropped pairs
    r1, c1 = small.pair_results(), small.counters()
    small.close()
    assert c1["hull_overflow"] == c0["inside"] - 2
    dropped = r1["status"] == PAIR_DROPPED
    assert dropped.sum() == c1["hull_overflow"] and ((r0["status"][dropped] & 0xFF) == 2).all()
    keep = ~dropped & ((r0["status"] & 0xFF) != 0)
    assert np.array_equal(r1["status"][keep], r0["status"][keep]) and np.array_equal(r1["planes"][keep], r0["planes"][keep])


def test_coplanar_cloud_is_flagged_degenerate(golden, product_lib):
    """qhull would merge coplanar facets (C-0); the replay does not restate merging and must say so"""
    g = np.stack(np.meshgrid(np.arange(4.0), np.arange(4.0), np.arange(4.0), indexing="ij"), -1).reshape(-1, 3)  # cube lattice
    ctx = make_ctx(golden, n_agents=6, max_neighbors=5, horizon=20)
    try:
        planes, verts, deg = ctx.convex_hull(g)
        assert deg != 0
    except LqroError:
        pass  # "hull failed (flat input or internal limit)" is the other legal answer
    flat = np.c_[np.random.default_rng(0).normal(size=(50, 2)), np.zeros(50)]
    with pytest.raises(LqroError):
        ctx.convex_hull(flat)
    ctx.close()


def test_size_limits_fail_loudly(golden, product_lib):
    from lqr_obstacles_b200 import Lqro

    with pytest.raises(LqroError):
        Lqro(n_agents=4, max_neighbors=3, horizon=400)          # H*P beyond the shared-memory budget
    with pytest.raises(LqroError):
        Lqro(n_agents=0, max_neighbors=3, horizon=20)
    ctx = Lqro(n_agents=4, max_neighbors=3, horizon=20)
    with pytest.raises(LqroError):                               # no dynamics yet
        ctx.step(np.zeros((4, 16)), np.zeros((4, 3)), np.zeros((4, 3), dtype=np.int32))
    ctx.close()


def test_maximum_reachable_cloud(golden, oracle, product_lib):
    """two agents on top of each other, zero relative velocity: every one of the H*P candidates is reachable-tested and
    the cloud overflows the 256-point shared part (overflow path in HBM); reference horizon H = 30"""
    x = np.zeros((2, 16))
    x[:, 12:] = 9.80665 * 0.5 / 4
    x[1, :3] = [0.05, -0.02, 0.01]
    nbr = np.array([[1], [0]], dtype=np.int32)
    vp = np.zeros((2, 3))
    ctx = make_ctx(golden, n_agents=2, max_neighbors=1, horizon=30)
    v = ctx.step(x, vp, nbr)
    res = ctx.pair_results()
    NF, T, sph = ctx.tables()
    ov, opl, ost, _ = oracle.step(NF, T, sph, x, vp, nbr)
    assert res["nreach"].max() > 256
    assert np.array_equal(res["status"] & 0xFF, ost & 0xFF) and np.array_equal(v, ov)
    st = ost & 0xFF
    assert np.array_equal(res["planes"][st != 0], opl[st != 0])
    ctx.close()


def test_chunked_neighbour_upload_equals_plain(golden, product_lib, monkeypatch):
    """LQRO_FLAG_NO_OVERLAP: host-buffer steps of >= 2^20 pairs upload the neighbour table in blocks overlapped with the pair
    kernel (one launch per block): same planes, statuses and velocities as the single-launch path fed with a resident table --
    and as the default step of a large swarm (polling hull launch as a concurrent consumer next to the pair kernel)"""
    from lqr_obstacles_b200.api import FLAG_NO_OVERLAP

    x, vp = swarm.lattice_swarm((32, 32, 16))
    N, K = len(x), 64
    nbr = swarm.knn(x[:, :3], K)
    a = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20, flags=FLAG_NO_OVERLAP)
    v1 = a.step(x, vp, nbr).copy()          # chunked (N*K = 2^20)
    r1 = a.pair_results()
    assert a.counters()["launches"] > 4
    v2 = a.step(x, vp, None).copy()         # table already resident: one pair-kernel launch
    r2 = a.pair_results()
    assert a.counters()["launches"] == 4
    assert np.array_equal(v1, v2) and np.array_equal(r1["status"], r2["status"]) and np.array_equal(r1["nreach"], r2["nreach"])
    monkeypatch.setenv("LQRO_NBR_STREAM_MIN_PAIRS", str(1 << 17))   # as a step of >= 2^21 pairs: the table goes up next to the pair kernel
    b = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20)
    monkeypatch.delenv("LQRO_NBR_STREAM_MIN_PAIRS")
    v3 = b.step(x, vp, nbr).copy()          # default: polling hull launch + pair kernel + wide hull launch + LP
    r3 = b.pair_results()
    assert b.counters()["launches"] == 4 and b.last_step_nbr_streamed()   # the table went up next to the pair kernel
    assert np.array_equal(v3, v2) and np.array_equal(r3["status"], r2["status"]) and np.array_equal(r3["nreach"], r2["nreach"])
    has3 = (r3["status"] & 0xFF) != 0
    assert np.array_equal(r3["planes"][has3], r2["planes"][has3])
    b.close()
    c = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20)   # below 2^21 pairs: the table is copied in front of the pair kernel
    v4 = c.step(x, vp, nbr).copy()
    r4 = c.pair_results()
    assert not c.last_step_nbr_streamed()
    assert np.array_equal(v4, v2) and np.array_equal(r4["status"], r2["status"]) and np.array_equal(r4["planes"][has3], r2["planes"][has3])
    c.close()
    has = (r1["status"] & 0xFF) != 0
    assert np.array_equal(r1["planes"][has], r2["planes"][has])
    v3 = a.step(x, vp, nbr)                 # and again chunked, after a plain step
    assert np.array_equal(v3, v1)
    a.close()


@pytest.mark.parametrize("H,P", [(12, 32), (25, 40), (7, 64), (20, 50), (30, 50)])
def test_other_horizons_and_sample_counts(golden, oracle, product_lib, H, P):
    """(H, P) = (20, 50) and (30, 50) run kernels specialised at compile time, everything else the generic ones:
    each against the oracle with the same H and P, hull branch included"""
    x, vp = _tight_swarm()
    N, K = len(x), 10
    nbr = swarm.knn(x[:, :3], K)
    ctx = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=H, n_samples=P)
    v = ctx.step(x, vp, nbr)
    res = ctx.pair_results()
    NF, T, sph = ctx.tables()
    oNF, oT, _ = oracle.horizon(golden["dyn_A"], golden["dyn_B"], golden["dyn_L"], golden["dyn_E"], H)
    assert np.array_equal(T, oT) and np.array_equal(sph, oracle.spheres(P))
    ov, opl, ost, _ = oracle.step(NF, T, sph, x, vp, nbr, nthreads=4)
    st = res["status"] & 0xFF
    assert np.array_equal(st, ost & 0xFF)
    assert np.array_equal(res["planes"][st != 0], opl[st != 0])
    assert np.array_equal(v, ov)
    assert (st == 2).sum() >= 1 or H < 20   # short horizons see no inside pair in this swarm
    ctx.close()


def test_slot_major_enumeration_widths(golden, product_lib, monkeypatch):
    """the order in which the pair kernel hands the pairs out (slot-major chunks of width first, first, 2 first, ... -- powers of two
    take a shift, anything else a division, lqro_step.cuh) does not enter any result: LQRO_FIRST_SLOTS = 3, 5 and 32 against
    the default, on a swarm large enough for the polling hull launch (>= 2^17 pairs) and on a small one"""
    for shape, K in (((16, 16, 16), 48), ((6, 6, 4), 24)):
        x, vp = swarm.lattice_swarm(shape)
        x[:, :3] *= 0.5  # as _tight_swarm: some pairs take the hull branch
        nbr = swarm.knn(x[:, :3], K)
        want = None
        for first in (None, 3, 5, 32):
            if first is None:
                monkeypatch.delenv("LQRO_FIRST_SLOTS", raising=False)
            else:
                monkeypatch.setenv("LQRO_FIRST_SLOTS", str(first))
            ctx = make_ctx(golden, n_agents=len(x), max_neighbors=K, horizon=20)
            v = ctx.step(x, vp, nbr)
            res = ctx.pair_results()
            got = (v, res["status"], res["planes"], res["nreach"])
            ctx.close()
            if want is None:
                want = got
                assert ((want[1] & 0xFF) == 2).sum() > 0  # hulls among them
            else:
                for a, b in zip(got, want):
                    assert np.array_equal(a, b), (shape, first)
