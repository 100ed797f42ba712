"""Scenario data for the path: the reference's start/goal sets and the synthetic swarms of
BASELINE.json (configs 4 and 5).  Host-side input generation only (numpy); nothing here computes
any part of the hot path.

Synthetic swarm: jittered lattice, spacing 4 m, jitter U(-1,1) m per axis; velocity
N(0,1) m/s per axis clipped to +-3; x = [p, v, 0, 0, F_hover]; preferred velocity 1.5 m/s toward the
lattice-antipodal cell; neighbours = K nearest by squared distance (ties to the lower index, the
insertAgentNeighbor rule, AGENT:153-174).  Random numbers are splitmix64(seed + stream*2^32 + index),
so a C++ or CUDA generator reproduces the same data.
"""
import numpy as np

SEED = 20261018
GRAVITY, MASS = 9.80665, 0.5  # MAIN:179,186
F_HOVER = GRAVITY * MASS / 4  # nominalInput, MAIN:197

MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    with np.errstate(over="ignore"):
        z = (np.asarray(z, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & MASK
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & MASK
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & MASK
        return z ^ (z >> np.uint64(31))


def uniform01(stream, index, seed=SEED):
    with np.errstate(over="ignore"):
        key = np.uint64(seed) + (np.uint64(stream) << np.uint64(32)) + np.asarray(index, dtype=np.uint64)
        bits = splitmix64(key)
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal01(stream, index, seed=SEED):
    u1 = uniform01(2 * stream, index, seed)
    u2 = uniform01(2 * stream + 1, index, seed)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def lattice_swarm(dims, spacing=4.0, seed=SEED):
    """-> x (N,16) state estimates, vpref (N,3)."""
    nx, ny, nz = dims
    n = nx * ny * nz
    idx = np.arange(n, dtype=np.uint64)
    ix, iy, iz = np.unravel_index(np.arange(n), dims)
    cell = np.stack([ix, iy, iz], axis=1).astype(np.float64)
    centre = (np.array(dims) - 1) / 2.0
    pos = (cell - centre) * spacing
    jit = np.stack([uniform01(s, idx, seed) for s in (0, 1, 2)], axis=1) * 2.0 - 1.0
    pos = pos + jit
    vel = np.clip(np.stack([normal01(s, idx, seed) for s in (3, 4, 5)], axis=1), -3.0, 3.0)
    x = np.zeros((n, 16))
    x[:, 0:3] = pos
    x[:, 3:6] = vel
    x[:, 12:16] = F_HOVER
    anti = -(cell - centre) * spacing  # centre of the antipodal cell
    d = anti - pos
    nrm = np.linalg.norm(d, axis=1, keepdims=True)
    vpref = np.where(nrm > 1e-9, 1.5 * d / np.maximum(nrm, 1e-9), 0.0)
    return np.ascontiguousarray(x), np.ascontiguousarray(vpref)


def knn(pos, K):
    """K nearest other agents by squared distance, ties to the lower index. -> (N,K) int32 (-1 = none)."""
    from scipy.spatial import cKDTree

    n = len(pos)
    k = min(K + 1, n)
    dist, idx = cKDTree(pos).query(pos, k=k)
    if k == 1:
        dist, idx = dist[:, None], idx[:, None]
    srt = np.lexsort((idx, dist), axis=1)  # ascending distance, equal distances by index
    idx_s = np.take_along_axis(idx, srt, axis=1)
    notself = idx_s != np.arange(n)[:, None]
    slot = np.cumsum(notself, axis=1) - 1  # compact the non-self entries to the left
    sel = notself & (slot < K)
    rows = np.broadcast_to(np.arange(n)[:, None], (n, k))
    out = np.full((n, K), -1, dtype=np.int32)
    out[rows[sel], slot[sel]] = idx_s[sel].astype(np.int32)
    return out


def all_pairs_neighbors(n):
    """The reference's neighbour set: every other agent, in index order (MAIN:1376-1380)."""
    out = np.zeros((n, n - 1), dtype=np.int32)
    for i in range(n):
        out[i] = [j for j in range(n) if j != i]
    return out


def hexagon6():
    """MAIN:1163-1218: hexagon of radius 6, antipodal goals."""
    d = np.sqrt(3.0)
    pts = np.array([[6.0, 0, 0], [3.0, 3.0 * d, 0], [-3.0, 3.0 * d, 0], [-6.0, 0, 0], [-3.0, -3.0 * d, 0], [3.0, -3.0 * d, 0]])
    return pts.copy(), pts[[3, 4, 5, 0, 1, 2]].copy()


def quad24():
    """The 24-quad set of 'LQRObstacles (visionsystem).cpp':1313-1460 (also the commented block
    MAIN:1220-1330); its first 4/8/16 entries serve as the 4/8/16-quad scenarios."""
    a, a2, a3 = np.sqrt(8.0), np.sqrt(18.0), np.sqrt(2.0)
    s = [(6, 0), (-6, 0), (0, 6), (0, -6), (4, 0), (-4, 0), (0, 4), (0, -4), (a, a), (a, -a), (-a, -a), (-a, a), (a2, a2), (a2, -a2),
         (-a2, -a2), (-a2, a2), (2, 0), (-2, 0), (0, 2), (0, -2), (a3, a3), (a3, -a3), (-a3, -a3), (-a3, a3)]
    g = [(-6, 0), (6, 0), (0, -6), (0, 6), (-4, 0), (4, 0), (0, -4), (0, 4), (-a2, -a2), (-a2, a2), (a2, a2), (a2, -a2), (-a, -a), (-a, a),
         (a, a), (a, -a), (-8, 0), (8, 0), (0, -8), (0, 8), (-a3, -a3), (-a3, a3), (a3, a3), (a3, -a3)]
    S = np.array([[x, y, 0.0] for x, y in s])
    G = np.array([[x, y, 0.0] for x, y in g])
    return S, G


def circle_antipodal(n, radius):
    """32/64-quad initial conditions are not in the reference: circle of `radius`, antipodal goals."""
    th = 2 * np.pi * np.arange(n) / n
    S = np.stack([radius * np.cos(th), radius * np.sin(th), np.zeros(n)], axis=1)
    return S, -S
