"""CPU tests: min-snap planner (host) and the reference-generation oracle against outputs of the
reference's own code (tests/golden/refgen_golden.npz, tests/golden/make_refgen_golden.py)."""
import numpy as np
import pytest

from conftest import golden
from ndp_nmpc_qd_b200 import traj_gen
from oracle import refgen_numpy as orf

NAMES = ["eight_high_dyn", "eight_low", "eight_low_diff_h"]


@pytest.mark.parametrize("name", NAMES)
def test_planner_matches_reference_coefficients(name):
    g = golden("refgen_golden.npz")
    p = traj_gen.PATHS[name]
    assert np.array_equal(p["pos"].T, g[name + "_wpts"]) and np.allclose(np.radians(p["yaw_deg"]), g[name + "_yaw"])
    tr = traj_gen.plan_named(name)
    assert np.allclose(tr.t_cum, g[name + "_t_cum"], rtol=1e-14)
    for mine, key in ((tr.cx, "_cx"), (tr.cy, "_cy"), (tr.cz, "_cz"), (tr.cyaw, "_cyaw")):
        ref = g[name + key].reshape(mine.shape)
        assert np.abs(mine - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_points(name):
    g = golden("refgen_golden.npz")
    m = len(g[name + "_t_cum"]) - 1
    tr = traj_gen.Trajectory(g[name + "_t_cum"], g[name + "_cx"].reshape(m, 8), g[name + "_cy"].reshape(m, 8), g[name + "_cz"].reshape(m, 8),
                             g[name + "_cyaw"].reshape(m, 4), g[name + "_wpts"][:, -1])
    for t, x, u in zip(g[name + "_t"], g[name + "_x"], g[name + "_u"]):
        xo, uo = orf.ref_point(tr, float(t))
        assert np.abs(xo - x).max() < 1e-11 and np.abs(uo - u).max() < 1e-10, t


def test_survey_trajectory_characteristics():
    """duration 23.131 s, v_max 9.62 m/s, c in [9.64, 20.33] on eight_high_dyn."""
    tr = traj_gen.plan_named("eight_high_dyn")
    assert abs(tr.duration - 23.131) < 1e-3
    ts = np.arange(0, tr.duration, 0.02)
    pts = [orf.ref_point(tr, float(t)) for t in ts]
    v = np.array([np.linalg.norm(p[0][3:6]) for p in pts]); c = np.array([p[1][3] for p in pts])
    assert abs(v.max() - 9.617) < 0.01 and abs(c.min() - 9.638) < 0.01 and abs(c.max() - 20.327) < 0.01
