"""CPU checks of the boundary: the C-ABI library loads, exports every symbol include/lqro.h declares,
and refuses to run without a GPU (no fallback).  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lqro.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lqro_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(product_lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(product_lib, s), "liblqro.so does not export %s" % s


def test_python_mirror_lists_every_symbol(product_lib):
    from lqr_obstacles_b200 import api

    assert sorted(api.EXPORTS) == declared_symbols()


def test_default_config_is_the_reference_constants(product_lib):
    from lqr_obstacles_b200 import default_config

    c = default_config()
    # MAIN:30-36, 1102, 1367
    assert (c.n_agents, c.horizon, c.n_samples) == (6, 30, 50)
    assert (c.xy_radius, c.z_radius, c.max_speed, c.lp_radius) == (0.26, 0.75, 30.0, 100.0)
    assert c.hull_input == 0 and c.fix_facet0_normal == 1


def test_no_cpu_fallback(product_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lqr_obstacles_b200 import Lqro, LqroError

    with pytest.raises(LqroError, match="no CPU fallback"):
        Lqro()


def test_product_does_not_link_the_oracle(product_lib):
    """The product library must not depend on anything under oracle/."""
    so = os.path.join(ROOT, "lqr_obstacles_b200", "liblqro.so")
    blob = open(so, "rb").read()
    assert b"liblqro_oracle" not in blob and b"liblqro_ref" not in blob
    for f in os.listdir(os.path.join(ROOT, "lqr_obstacles_b200")):
        if f.endswith(".py"):
            txt = open(os.path.join(ROOT, "lqr_obstacles_b200", f)).read()
            assert "oracle" not in txt.replace("no CPU fallback", ""), f


def test_bad_arguments_return_codes(product_lib):
    from lqr_obstacles_b200.api import Config

    c = Config()
    product_lib.lqro_default_config(ctypes.byref(c))
    c.horizon = 0
    h = ctypes.c_void_p()
    assert product_lib.lqro_create(ctypes.byref(c), ctypes.byref(h)) == -1  # LQRO_E_ARG
    assert product_lib.lqro_create(None, ctypes.byref(h)) == -1
    assert product_lib.lqro_destroy(None) == 0
    # simulator-form update: no context / no buffers => LQRO_E_ARG, never a crash
    assert product_lib.lqro_sim_update(None, 1, *([None] * 10)) == -1
    assert product_lib.lqro_sim_update_device(None, 1, *([None] * 11)) == -1
    assert product_lib.lqro_sim_set_model(None, None) == -1
    seed = ctypes.c_uint(1)
    assert product_lib.lqro_msvc_normals(ctypes.byref(seed), -1, None) == -1
    assert product_lib.lqro_msvc_normals(None, 1, None) == -1
    assert product_lib.lqro_msvc_normals(ctypes.byref(seed), 0, None) == 0 and seed.value == 1


def test_headless_simulator_refuses_to_run_without_a_gpu(product_lib):
    """tools/headless/lqro_headless is product code: with no CUDA device it must fail loudly, not fall back"""
    import subprocess

    import torch
    from lqr_obstacles_b200 import build

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    exe = build.build_headless()
    r = subprocess.run([exe, "--steps", "1", "--quiet"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


def test_swarm_generator_is_deterministic():
    from lqr_obstacles_b200 import swarm

    x1, v1 = swarm.lattice_swarm((4, 4, 4))
    x2, v2 = swarm.lattice_swarm((4, 4, 4))
    assert np.array_equal(x1, x2) and np.array_equal(v1, v2)
    assert x1.shape == (64, 16) and np.allclose(np.linalg.norm(v1, axis=1), 1.5)
    assert np.all(np.abs(x1[:, 3:6]) <= 3.0) and np.all(x1[:, 12:] == swarm.F_HOVER)
    # splitmix64 known answer (first output of the reference implementation for state 0)
    assert int(swarm.splitmix64(np.uint64(0))) == 0xE220A8397B1DCDAF
    nb = swarm.knn(x1[:, :3], 8)
    assert nb.shape == (64, 8) and (nb >= 0).all()
    for i in (0, 17, 63):
        d = np.linalg.norm(x1[:, :3] - x1[i, :3], axis=1)
        d[i] = np.inf
        assert list(nb[i]) == list(np.argsort(d, kind="stable")[:8])
    assert swarm.all_pairs_neighbors(4).tolist() == [[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]]
