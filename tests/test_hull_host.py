"""CPU check of the CUDA hull's LOGIC: the kernel source csrc/lqro_hull_core.cuh compiled as plain C++ on a 32-thread
lockstep emulation of the warp primitives (tests/hull_host) against oracle/hull_cpu.cpp -- facet list, facet order,
vertex order and plane bits -- on the reference's own qhull fixture, on inside-pair clouds of the path and on
generated clouds that drive the rare branches (large cones, exhaustive partition, horizon flood, ring compaction, pool
exhaustion).  The GPU parity tests (tests/test_gpu_parity.py) run the same source on the device."""
import numpy as np
import pytest

from _libs import HullHost


@pytest.fixture(scope="module")
def host():
    return HullHost()


def _same_hull(host, oracle, pts, mode=0, cap_f=0):
    """one warp per hull AND two warps per hull (main + helper with the speculative horizon search, as k_hull runs it)"""
    cloud = oracle.round6(pts) if mode else pts
    oplanes, overts, order, info = oracle.hull(cloud)
    for coop in (False, True):
        planes, verts, hi, _ = host.hull(pts, mode=mode, cap_f=cap_f, coop=coop)
        assert hi["fail"] == 0
        assert np.array_equal(verts, overts)     # facet list, order and vertex order
        assert np.array_equal(planes, oplanes)   # same formula, no FMA: identical bits
        assert hi["ninserted"] == info["ninserted"]
        assert hi["degenerate"] == (1 if info["degenerate"] else 0)
    return len(overts)


def test_round6_is_printf_g(host):
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.normal(size=300) * 10.0 ** rng.integers(-8, 8, size=300),
                        [0.0, 1.0, -2.5, 1.015625, 123456.5, 999999.5, 1e-5, 99999.95, 0.0999999949, 1234565.0, -31.41585]])
    assert np.array_equal(host.round6(v), np.array([float("%g" % x) for x in v]))


def test_reference_qhull_fixture(host, golden):
    """pointList.txt -> Planes.txt + facetVertices.txt of the reference tree: same 62 facets, order, vertex order"""
    planes, verts, hi, _ = host.hull(golden["qh_points"])
    assert hi["fail"] == 0 and hi["degenerate"] == 0
    assert np.array_equal(verts, golden["qh_facets"])
    assert np.abs(planes - golden["qh_planes"]).max() < 1e-13


def test_inside_pair_clouds_of_the_path(host, oracle, golden):
    """reachable clouds of inside pairs (tight 4x4x4 lattice, H = 20), rendered to 6 digits as MAIN:765-767 does"""
    from lqr_obstacles_b200 import swarm

    x, _ = swarm.lattice_swarm((4, 4, 4))
    x[:, 0:3] *= 0.5
    nbr = swarm.knn(x[:, :3], 3)
    NF, T, _ = oracle.horizon(golden["dyn_A"], golden["dyn_B"], golden["dyn_L"], golden["dyn_E"], 20)
    sph = oracle.spheres(50)
    done = 0
    for i in range(len(x)):
        for j in nbr[i]:
            r = oracle.pair(NF, T, sph, x[i], x[j])
            if r["status"] != 2:
                continue
            q = oracle.candidates(NF, T, sph, x[i], x[j])
            pts = q[oracle.reachable(q, x[i], x[j])]
            nf = _same_hull(host, oracle, pts, mode=1)
            assert nf == r["n_facets"]
            vrel = x[i, 3:6] - x[j, 3:6]
            _, _, hi, esc = host.hull(pts, mode=1, vrel=vrel)
            assert hi["best_facet"] == r["best_facet"]
            assert abs(esc[0] - r["distance"]) <= 1e-12 * max(1.0, abs(r["distance"]))
            assert np.abs(esc[1] - r["normal"]).max() <= 1e-15
            done += 1
            if done >= 8:
                return
    assert done >= 4


def _clouds(rng):
    p = rng.normal(size=(300, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    yield p, 0                                                   # every point a vertex
    yield rng.normal(size=(1500, 3)), 0                          # long outside lists: multi-chunk partitions
    yield rng.uniform(-1, 1, size=(700, 3)), 0                   # box: sharp cones -> exhaustive partition mode
    yield rng.normal(size=(400, 3)) * [30, 3, 0.3], 1            # anisotropic, rendered to 6 digits
    yield np.round(rng.uniform(-1, 1, size=(130, 3)), 2), 0      # coarse grid: neighbours within SEARCHdist (horizon flood)
    for n in (4, 5, 9):
        yield rng.normal(size=(n, 3)), 0


def test_generated_clouds_and_rare_branches(host, oracle):
    host.coverage(reset=True)
    for p, mode in _clouds(np.random.default_rng(5)):
        try:
            _same_hull(host, oracle, p, mode=mode)
        except RuntimeError:   # the oracle refuses a degenerate cloud
            pass
    cov = host.coverage()
    assert cov["multi_chunk_partition"] > 0 and cov["exhaustive"] > 0, cov
    # the outside sets live in contiguous segments: their storage was compacted, a facet that already held points received
    # more (segment moved), and long partitions went through the cold arrays
    assert cov["olist_compaction"] > 0 and cov["append_moved_segment"] > 0 and cov["gather_cold"] > 0, cov
    # two warps per hull: the helper's speculative horizon was adopted, rejected (new furthest point) and unavailable (empty ring)
    assert cov["spec_adopted"] > 0 and cov["spec_none"] > 0, cov


def test_general_branches_forced(oracle):
    """the build that sends every cone of more than 2 facets through the owner-word ridge matching and every walk over
    more than 3 cone facets through the wide visited set (the branches large cones take on the device)"""
    general = HullHost(general=True)
    general.coverage(reset=True)
    rng = np.random.default_rng(8)
    for n in (40, 150, 600):
        _same_hull(general, oracle, rng.normal(size=(n, 3)))
    p = rng.normal(size=(200, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    assert _same_hull(general, oracle, p) == 2 * 200 - 4
    cov = general.coverage()
    assert cov["cone_gt16"] > 0 and cov["cone_gt64"] > 0 and cov["append_sweep_buckets"] > 0 and cov["spec_rejected"] > 0, cov


def test_ring_compaction(host, oracle):
    """a pending ring that fills up with entries of facets deleted while pending is compacted in place"""
    host.coverage(reset=True)
    rng = np.random.default_rng(20)
    p = rng.normal(size=(380, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    oplanes, overts, _, _ = oracle.hull(p)
    planes, verts, hi, _ = host.hull(p, ring=128)
    assert hi["fail"] == 0 and np.array_equal(verts, overts) and np.array_equal(planes, oplanes)
    assert host.coverage()["ring_compaction"] > 0
    _, _, hi, _ = host.hull(p, ring=32)   # fewer ring entries than pending facets: reported as pool exhaustion
    assert hi["fail"] == 2


def test_pool_exhaustion_is_reported(host):
    """a facet pool that is too small must end with fail code 2 (the kernel then re-runs on the full-size pool)"""
    rng = np.random.default_rng(6)
    p = rng.normal(size=(400, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    _, _, hi, _ = host.hull(p, cap_f=64)
    assert hi["fail"] == 2 and hi["nfacets"] == 0


def test_flat_input_fails_cleanly(host):
    p = np.zeros((10, 3))
    p[:, 0] = np.arange(10)
    _, _, hi, _ = host.hull(p)
    assert hi["fail"] == 1
