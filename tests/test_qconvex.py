"""Wire formats: tools/qconvex/qconvex is a `qconvex` the unmodified reference can spawn
(MAIN:769-770); it must reproduce the reference's own fixture files (pointList.txt -> Planes.txt,
facetVertices.txt, committed to the golden vectors as qh_points / qh_planes / qh_facets)."""
import os
import subprocess

import numpy as np
import pytest

from lqr_obstacles_b200 import build as lbuild


def _write_pointlist(path, pts):
    # convexHullFiles, MAIN:764-768: "3\n<n>\n" then "x y z " per line, default ostream precision (6 significant digits)
    with open(path, "w") as f:
        f.write("3\n%d\n" % len(pts))
        for p in pts:
            f.write("%g %g %g \n" % tuple(p))


def test_qconvex_builds():
    exe = lbuild.build_qconvex()
    assert os.access(exe, os.X_OK)
    r = subprocess.run([exe, "bogus"], input="", capture_output=True, text=True)
    assert r.returncode == 2 and "unsupported option" in r.stderr


@pytest.mark.gpu
def test_qconvex_reproduces_reference_fixture(golden, product_lib, tmp_path):
    exe = lbuild.build_qconvex()
    pl = tmp_path / "pointList.txt"
    _write_pointlist(pl, golden["qh_points"])
    for opt, out in (("n", "Planes.txt"), ("Fv", "facetVertices.txt")):
        with open(pl) as fin:
            subprocess.run([exe, opt, "TO", str(tmp_path / out)], stdin=fin, check=True)
    # parse the way the reference does (MAIN:776-803)
    rows = open(tmp_path / "Planes.txt").read().split("\n")
    assert rows[0] == "4" and int(rows[1]) == 62
    planes = np.array([[float(v) for v in r.split()] for r in rows[2:2 + 62]])
    assert np.abs(planes - golden["qh_planes"]).max() < 1e-13
    rows = open(tmp_path / "facetVertices.txt").read().split("\n")
    assert int(rows[0]) == 62
    fv = np.array([[int(v) for v in r.split()] for r in rows[1:1 + 62]])
    assert (fv[:, 0] == 3).all() and np.array_equal(fv[:, 1:], golden["qh_facets"])
    # stdout form (no TO)
    with open(pl) as fin:
        r = subprocess.run([exe, "Fv"], stdin=fin, capture_output=True, text=True, check=True)
    assert r.stdout == open(tmp_path / "facetVertices.txt").read()


@pytest.mark.gpu
def test_unmodified_reference_spawns_gpu_qconvex(golden, reference, product_lib, monkeypatch):
    """The reference's convexHull (MAIN:809-838), compiled unmodified, with its system("qconvex ...") calls going
    to the real shell: tools/qconvex/qconvex on PATH answers them from the GPU.  Same distance / normal as with
    the in-process CPU stand-in (the oracle's qhull replay), bit for bit."""
    exe = lbuild.build_qconvex()
    monkeypatch.setenv("PATH", os.path.dirname(exe) + os.pathsep + os.environ.get("PATH", ""))
    rng = np.random.default_rng(11)
    for trial in range(3):
        pts = rng.normal(size=(300 + 100 * trial, 3)) * [3.0, 2.0, 1.0]
        vrel = rng.normal(size=3) * 0.2
        d_cpu, n_cpu = reference.convex_hull(vrel, pts)
        reference.lib.ref_set_system_passthrough(1)
        try:
            d_gpu, n_gpu = reference.convex_hull(vrel, pts)
        finally:
            reference.lib.ref_set_system_passthrough(0)
        assert d_gpu == d_cpu and np.array_equal(n_gpu, n_cpu)
