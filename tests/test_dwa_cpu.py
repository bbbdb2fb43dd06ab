"""CPU: the DynamicWindow restatement (oracle/ergodic_oracle.c) against the compiled reference
(DynamicWindow::control, both overloads, built from the unmodified dynamic_window.cpp)."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle, RefLib
from test_collision_cpu import random_map

needs_ref = pytest.mark.skipif(not RefLib.available(), reason="compiled reference (oracle/_ref) not present")

# explore_omni.yaml:65-70 + exploration_omni_node.cpp: dt, horizon, acc_dt, acc limits, velocity limits
DWA_OMNI = (0.1, 2.0, 0.2, 1.0, 1.0, 2.0, 1.0, -1.0, 1.0, -1.0, 2.0, -2.0)
DWA_CART = (0.1, 2.0, 0.2, 1.0, 0.0, 2.0, 1.0, -1.0, 0.0, 0.0, 2.0, -2.0)   # exploration_cart_node.cpp:198
COL = (0.2, 1.0, 0.05, 0.65)


def scenario(rng, n, ys=160, xs=200, res=0.05, p_occ=0.004):
    data = random_map(rng, ys, xs, p_occ=p_occ)
    x0 = np.column_stack([rng.uniform(0.3, xs * res - 0.3, n), rng.uniform(0.3, ys * res - 0.3, n),
                          rng.uniform(-np.pi, np.pi, n)])
    vb = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-2, 2, n)])
    return data, res, x0, vb


@needs_ref
@pytest.mark.parametrize("cfg,samples", [(DWA_OMNI, (3, 8, 5)), (DWA_CART, (5, 1, 9)), (DWA_OMNI, (0, 2, 1))])
def test_dwa_twist_reference_matches(cfg, samples):
    rng = np.random.default_rng(sum(samples))
    data, res, x0, vb = scenario(rng, 250)
    vref = np.column_stack([rng.uniform(-1, 1, 250), rng.uniform(-1, 1, 250), rng.uniform(-2, 2, 250)])
    fo, uo, _ = Oracle.dwa_control(data, res, 0.0, 0.0, COL, cfg, samples, x0, vb, vref=vref)
    fr, ur = RefLib.dwa_control(data, res, 0.0, 0.0, COL, cfg, samples, x0, vb, vref=vref)
    np.testing.assert_array_equal(fo, fr)
    np.testing.assert_array_equal(uo, ur)
    assert 0 < fo.sum() <= len(fo)


@needs_ref
@pytest.mark.parametrize("cfg,samples", [(DWA_OMNI, (3, 8, 5)), (DWA_CART, (4, 1, 7))])
def test_dwa_trajectory_reference_matches(cfg, samples):
    rng = np.random.default_rng(17 + sum(samples))
    data, res, x0, vb = scenario(rng, 200)
    t = np.arange(50) * 0.1
    xt_ref = np.column_stack([4.0 + 1.5 * np.cos(0.7 * t), 3.5 + 1.2 * np.sin(0.9 * t), 4.0 * np.sin(0.5 * t)])  # yaw beyond pi
    fo, uo, _ = Oracle.dwa_control(data, res, 0.0, 0.0, COL, cfg, samples, x0, vb, xt_ref=xt_ref, dt_ref=0.1)
    fr, ur = RefLib.dwa_control(data, res, 0.0, 0.0, COL, cfg, samples, x0, vb, xt_ref=xt_ref, dt_ref=0.1)
    np.testing.assert_array_equal(fo, fr)
    np.testing.assert_array_equal(uo, ur)


@needs_ref
def test_dwa_no_solution():
    """every cell occupied: no collision-free twist, found = 0 and u_opt = 0 (dynamic_window.cpp:132-136)"""
    data = np.full((40, 40), 100, dtype=np.int8)
    x0, vb = np.array([[1.0, 1.0, 0.2]]), np.zeros((1, 3))
    fo, uo, cost = Oracle.dwa_control(data, 0.05, 0.0, 0.0, COL, DWA_OMNI, (3, 8, 5), x0, vb, vref=np.zeros((1, 3)))
    fr, ur = RefLib.dwa_control(data, 0.05, 0.0, 0.0, COL, DWA_OMNI, (3, 8, 5), x0, vb, vref=np.zeros((1, 3)))
    assert fo[0] == 0 and fr[0] == 0 and not uo.any() and not ur.any() and cost[0] > 1e300
