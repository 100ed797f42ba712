"""Scenario files and trajectory dumps (lqr_obstacles_b200/formats.py; the formats tools/headless/lqro_headless reads / writes)."""
import numpy as np
import pytest

from lqr_obstacles_b200 import formats, swarm


def test_scenario_round_trip(tmp_path):
    starts, goals = swarm.circle_antipodal(32, 12.0)
    p = str(tmp_path / "circle32.txt")
    formats.write_scenario(p, starts, goals)
    s2, g2 = formats.read_scenario(p)
    assert np.array_equal(s2, starts) and np.array_equal(g2, goals)  # repr() round-trips doubles exactly
    assert open(p).readline().strip() == "32"
    with pytest.raises(ValueError):
        formats.write_scenario(p, starts[:1], goals[:1])


def test_trajectory_round_trip_and_validation(tmp_path):
    rng = np.random.default_rng(0)
    xt, xe = rng.normal(size=(7, 5, 16)), rng.normal(size=(7, 5, 16))
    p = str(tmp_path / "traj.bin")
    formats.write_trajectory(p, xt, xe)
    a, b = formats.read_trajectory(p)
    assert np.array_equal(a, xt) and np.array_equal(b, xe)
    raw = open(p, "rb").read()
    assert raw[:8] == b"LQROTRJ1" and len(raw) == 24 + 8 * 2 * 7 * 5 * 16
    open(p, "wb").write(raw[:-8])
    with pytest.raises(ValueError):
        formats.read_trajectory(p)
    open(p, "wb").write(b"NOTATRAJ" + raw[8:])
    with pytest.raises(ValueError):
        formats.read_trajectory(p)
