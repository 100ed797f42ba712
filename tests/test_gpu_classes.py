"""Heterogeneous dynamics: agents in several dynamics classes, each class with its own gains
(L, E from the batched Riccati solve with different weights).  The obstacle of pair (i, j) uses agent i's class
(findFG(A, B, qlist[i]->L, qlist[i]->E, ...), MAIN:1384).  Expected values: the oracle run once per class on the
agents of that class, merged."""
import numpy as np
import pytest

from conftest import make_ctx
from lqr_obstacles_b200 import LqroError, swarm

pytestmark = pytest.mark.gpu


def _class_gains(oracle, golden):
    A, B = golden["dyn_A"], golden["dyn_B"]
    out = [(golden["dyn_L"], golden["dyn_E"])]
    for qv, qp, r in ((60.0, 0.1, 8.0), (150.0, 0.2, 3.0)):
        L, E, _, _ = oracle.riccati(A, B, qv, qp, r)
        out.append((L, E))
    return A, B, out


def _expected(oracle, A, B, gains, cls, x, vp, nbr, H):
    sph = oracle.spheres()
    N, K = nbr.shape
    newV = np.zeros((N, 3)); planes = np.zeros((N, K, 6), dtype=np.float32); status = np.zeros((N, K), dtype=np.int32)
    for c, (L, E) in enumerate(gains):
        sel = cls == c
        if not sel.any():
            continue
        NF, T, _ = oracle.horizon(A, B, L, E, H)
        sub = np.where(sel[:, None], nbr, -1).astype(np.int32)
        v, pl, st, _ = oracle.step(NF, T, sph, x, vp, sub, nthreads=4)
        newV[sel], planes[sel], status[sel] = v[sel], pl[sel], st[sel]
    return newV, planes, status


@pytest.mark.parametrize("nclass", [2, 3])
def test_classes_match_per_class_oracle(golden, oracle, product_lib, nclass):
    A, B, gains = _class_gains(oracle, golden)
    gains = gains[:nclass]
    x, vp = swarm.lattice_swarm((4, 4, 4))
    x[:, :3] *= 0.5  # some pairs take the hull branch
    N, K, H = len(x), 12, 20
    nbr = swarm.knn(x[:, :3], K)
    cls = np.random.default_rng(7).integers(0, nclass, N).astype(np.int32)
    ctx = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=H)
    for c, (L, E) in enumerate(gains):
        ctx.set_dynamics_class(c, A, B, L, E)
    ctx.set_agent_classes(cls)
    v = ctx.step(x, vp, nbr)
    res = ctx.pair_results()
    ev, epl, est = _expected(oracle, A, B, gains, cls, x, vp, nbr, H)
    st = res["status"] & 0xFF
    assert np.array_equal(st, est & 0xFF)
    assert (st == 2).sum() >= 1                       # the hull kernel switched tables per pair
    assert np.array_equal(res["planes"][st != 0], epl[st != 0])
    assert np.array_equal(v, ev)
    # the classes matter: a single-class step gives something else
    ctx.set_agent_classes(None)
    v0 = ctx.step(x, vp, nbr)
    assert not np.array_equal(v0, v)
    # and all agents in class 0 through the class path == the single-class path
    ctx.set_agent_classes(np.zeros(N, dtype=np.int32))
    assert np.array_equal(ctx.step(x, vp, nbr), v0)
    ctx.close()


def test_class_without_dynamics_is_an_error(golden, product_lib):
    ctx = make_ctx(golden, n_agents=8, max_neighbors=3, horizon=20)
    with pytest.raises(LqroError):
        ctx.set_agent_classes(np.array([0, 1, 0, 0, 0, 0, 0, 0], dtype=np.int32))
    with pytest.raises(LqroError):
        ctx.set_dynamics_class(16, golden["dyn_A"], golden["dyn_B"], golden["dyn_L"], golden["dyn_E"])
    ctx.close()
