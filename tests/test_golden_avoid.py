"""CPU: the collision / DynamicWindow restatement against the committed golden vectors generated
from the compiled reference (tests/golden/make_golden_avoid.py).  Flags and selected twists must
be identical."""
import os

import numpy as np

from oracle.pyoracle import Oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "avoid", "collision_dwa.npz")


def load():
    g = np.load(GOLD)
    args = (g["data"], float(g["res"]), float(g["xmin"]), float(g["ymin"]), tuple(g["col"]))
    return g, args


def test_collision_and_validate_against_golden():
    g, a = load()
    np.testing.assert_array_equal(Oracle.collision_check(*a, g["poses"]), g["hit"])
    np.testing.assert_array_equal(Oracle.validate_control(*a, g["inside"], g["twists"], 0.1, 0.5), g["valid_05"])
    np.testing.assert_array_equal(Oracle.validate_control(*a, g["inside"], g["twists"], 0.1, 2.0), g["valid_20"])
    assert 0 < g["hit"].sum() < len(g["hit"]) and 0 < g["valid_20"].sum() < len(g["valid_20"])


def test_dynamic_window_against_golden():
    g, a = load()
    cfg, smp = tuple(g["dwa_cfg"]), tuple(int(v) for v in g["samples"])
    f, u, _ = Oracle.dwa_control(*a, cfg, smp, g["inside"], g["twists"], vref=g["vref"])
    np.testing.assert_array_equal(f, g["dwa_found_twist"])
    np.testing.assert_array_equal(u, g["dwa_u_twist"])
    f, u, _ = Oracle.dwa_control(*a, cfg, smp, g["inside"], g["twists"], xt_ref=g["xt_ref"], dt_ref=0.1)
    np.testing.assert_array_equal(f, g["dwa_found_traj"])
    np.testing.assert_array_equal(u, g["dwa_u_traj"])
