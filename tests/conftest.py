import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))


@pytest.fixture(scope="session")
def oracle():
    from _libs import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own code (oracle/_ref).  Absent => the test is skipped (GPU box without a prebuilt _ref)."""
    from _libs import REF_SO, REFERENCE_TREE, Reference

    if not (os.path.exists(REF_SO) or os.path.isdir(REFERENCE_TREE)):
        pytest.skip("oracle/_ref not available here")
    return Reference()


@pytest.fixture(scope="session")
def product_lib():
    """liblqro.so must exist (built by __graft_entry__.build()); no fallback."""
    from lqr_obstacles_b200 import build, load_library

    build.build()
    return load_library()


def make_ctx(golden, **cfg):
    from lqr_obstacles_b200 import Lqro

    ctx = Lqro(**cfg)
    ctx.set_dynamics(golden["dyn_A"], golden["dyn_B"], golden["dyn_L"], golden["dyn_E"])
    return ctx


def state_pair_random(rng, u, mode):
    """random pair of agent states (same recipe as tests/golden/make_golden.py)"""
    xi, xj = np.zeros(16), np.zeros(16)
    xi[12:] = u
    xj[12:] = u
    d = rng.uniform(0.6, 3.0) if mode == 2 else (rng.uniform(3.0, 14.0) if mode == 1 else rng.uniform(25, 60))
    dirv = rng.normal(size=3)
    dirv /= np.linalg.norm(dirv)
    xi[:3], xj[:3] = dirv * d / 2 + rng.normal(size=3), -dirv * d / 2
    close = rng.uniform(0, 4)
    xi[3:6] = rng.normal(size=3) - dirv * close
    xj[3:6] = rng.normal(size=3) + dirv * close
    xi[6:12] = rng.normal(size=6) * 0.05
    xj[6:12] = rng.normal(size=6) * 0.05
    xi[12:] += rng.normal(size=4) * 0.05
    xj[12:] += rng.normal(size=4) * 0.05
    return xi, xj
