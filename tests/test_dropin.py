"""The drop-in boundary: include/lqro_dropin.hpp carries the reference's own signatures.
CPU: the demo translation unit compiles (syntax + template instantiation) with the stand-in types and, where
/root/reference exists, with the reference's OWN matrix.h / Vector3.h.  GPU: the program is linked against
liblqro.so, fed recorded states of the canonical 6-quad scenario, and its output compared with the reference's."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "dropin", "dropin_demo.cpp")
REF = "/root/reference"
CXX = "/usr/bin/g++"


def test_dropin_compiles_with_standins():
    subprocess.check_call([CXX, "-std=c++14", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), SRC])


def test_dropin_compiles_with_reference_types():
    if not os.path.isdir(os.path.join(REF, "include")):
        pytest.skip("reference tree not present")
    subprocess.check_call([CXX, "-std=c++14", "-fsyntax-only", "-fpermissive", "-w", "-DUSE_REFERENCE_HEADERS",
                           "-include", os.path.join(ROOT, "oracle", "shims", "pre.h"),
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(REF, "include"),
                           "-I" + os.path.join(REF, "QuadrotorHoverController"), SRC])


@pytest.mark.gpu
def test_dropin_program_matches_reference(golden, product_lib, tmp_path):
    exe = str(tmp_path / "dropin_demo")
    libdir = os.path.join(ROOT, "lqr_obstacles_b200")
    subprocess.check_call([CXX, "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"), SRC, "-o", exe, "-L" + libdir, "-llqro",
                           "-Wl,-rpath," + libdir])
    steps = list(range(0, 120, 8))
    lines = ["6"]
    for k in ("dyn_A", "dyn_B", "dyn_L", "dyn_E"):
        lines.append(" ".join(repr(float(v)) for v in golden[k].ravel()))
    for t in steps:
        for i in range(6):
            lines.append(" ".join(repr(float(v)) for v in golden["sc6_x_vgoal"][t][i]))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    it = iter([l for l in out if l.strip()])
    for t in steps:
        newv = np.array([[float(v) for v in next(it).split()] for _ in range(6)])
        shim = np.array([float(v) for v in next(it).split()[1:]])
        assert np.array_equal(newv, golden["sc6_newv"][t]), t        # lqro::Swarm::step == reference loop
        assert np.array_equal(shim, golden["sc6_newv"][t][0]), t     # per-function shims, agent 0
