"""Neighbour search: the device uniform-grid kNN against the oracle's literal restatement of
Agent::insertAgentNeighbor (AGENT:153-174), bit-exact index lists; and the oracle against an independent
kd-tree on the CPU."""
import numpy as np
import pytest

from conftest import make_ctx
from lqr_obstacles_b200 import swarm
from lqr_obstacles_b200.api import FLAG_DEVICE_KNN


def _states(pos):
    x = np.zeros((len(pos), 16))
    x[:, :3] = pos
    return x


def test_oracle_knn_vs_kdtree(oracle):
    """lqro_o_knn == scipy cKDTree + (distance, index) ordering on a jittered lattice (no exact ties expected)"""
    x, _ = swarm.lattice_swarm((8, 8, 8))
    a = oracle.knn(x, 16, nthreads=4)
    b = swarm.knn(x[:, :3], 16)
    assert np.array_equal(a, b)


def test_oracle_knn_ties_and_padding(oracle):
    """equal distances keep the lower index (strict < of AGENT:158,165); fewer than K others pad with -1"""
    g = np.stack(np.meshgrid(np.arange(3.0), np.arange(3.0), np.arange(3.0), indexing="ij"), -1).reshape(-1, 3)  # exact ties
    nb = oracle.knn(_states(g), 6)
    c = 13  # centre of the 3x3x3 cube: its 6 face neighbours tie at distance 1, ascending index
    assert list(nb[c]) == [4, 10, 12, 14, 16, 22]
    assert list(nb[0][:3]) == [1, 3, 9]
    few = oracle.knn(_states(g[:4]), 6)
    assert (few[:, 3:] == -1).all() and (few[:, :3] >= 0).all()
    for i in range(4):
        assert i not in few[i]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,K", [((16, 16, 16), 32), ((12, 10, 8), 64), ((9, 7, 5), 20), ((6, 6, 6), 100)])
def test_device_knn_vs_oracle(golden, oracle, product_lib, shape, K):
    x, _ = swarm.lattice_swarm(shape)
    ctx = make_ctx(golden, n_agents=len(x), max_neighbors=K, horizon=20)
    got = ctx.knn(x)
    want = oracle.knn(x, K, nthreads=8)
    assert np.array_equal(got, want)
    assert np.array_equal(ctx.neighbors(), want)
    ctx.close()


@pytest.mark.gpu
def test_device_knn_ties_flat_and_small(golden, oracle, product_lib):
    """integer lattice (massive exact ties), a planar swarm (zero z extent), coincident agents, N <= K"""
    g = np.stack(np.meshgrid(np.arange(7.0), np.arange(6.0), np.arange(5.0), indexing="ij"), -1).reshape(-1, 3)
    flat = g.copy(); flat[:, 2] = 0.25
    coinc = np.repeat(g[:40], 3, axis=0)
    six, _ = swarm.hexagon6()
    for pos, K in [(g, 32), (g, 64), (flat, 16), (coinc, 8), (six, 5), (six, 32), (g[:1], 4)]:
        x = _states(pos)
        ctx = make_ctx(golden, n_agents=len(x), max_neighbors=K, horizon=20)
        assert np.array_equal(ctx.knn(x), oracle.knn(x, K)), (len(pos), K)
        ctx.close()


@pytest.mark.gpu
def test_device_knn_sharded_rows_and_step(golden, oracle, product_lib):
    """a context that owns a block of agents lists only its rows; a step with nbr=None under LQRO_FLAG_DEVICE_KNN
    equals a step fed with the oracle's lists"""
    x, vp = swarm.lattice_swarm((8, 8, 4))
    N, K = len(x), 16
    want = oracle.knn(x, K, nthreads=4)
    a = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20, local_begin=64, n_local=96)
    assert np.array_equal(a.knn(x), want[64:160])
    a.close()
    b = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20, flags=FLAG_DEVICE_KNN)
    c = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20)
    v1 = b.step(x, vp, None)
    v2 = c.step(x, vp, want)
    assert np.array_equal(v1, v2)
    assert np.array_equal(b.neighbors(), want)
    b.close(); c.close()


@pytest.mark.gpu
def test_device_knn_full_size_sample(golden, oracle, product_lib):
    """bench workload (65,536 agents, K=64): 512 sampled rows against brute force; every row sorted, self-free"""
    x, _ = swarm.lattice_swarm((64, 32, 32))
    N, K = len(x), 64
    ctx = make_ctx(golden, n_agents=N, max_neighbors=K, horizon=20)
    got = ctx.knn(x)
    ctx.close()
    rows = np.random.default_rng(5).choice(N, 512, replace=False)
    for r in rows[:512]:
        assert np.array_equal(got[r], oracle.knn(x, K, begin=int(r), n_local=1)[0])
    assert (got >= 0).all() and (got != np.arange(N)[:, None]).all()
    d = ((x[got, :3] - x[:, None, :3]) ** 2).sum(-1)
    assert (np.diff(d, axis=1) >= -1e-9).all()
