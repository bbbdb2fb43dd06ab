"""CPU: the restatement (oracle/ergodic_oracle.c) against the golden vectors of the compiled reference for the
round-2 rows: Cart / Mecanum forward RK4 and model Jacobians, entropy of an occupancy grid; plus the reference's
own integrator known-answer test (test/test_integrator.cpp:44-73)."""
import os

import numpy as np

from oracle.pyoracle import Oracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_entropy.npz"))


def test_reference_integrator_kat():
    """test/test_integrator.cpp:44-73: Cart(0.1, 2.0), u = [1, 1], dt 0.1, 4 steps -> x = 0.01 (i + 1), y = theta = 0"""
    xt = G["kat_xt"]
    for i in range(4):
        assert abs(xt[i, 0] - 0.01 * (i + 1)) <= 4 * np.spacing(0.01 * (i + 1))  # ASSERT_DOUBLE_EQ: 4 ulp
        assert xt[i, 1] == 0.0 and xt[i, 2] == 0.0
    mine = Oracle.rk4_forward_cart(0.1, 2.0, 0.1, 0.4, [0.0, 0.0, 0.0], np.ones((4, 2)))
    assert np.array_equal(mine, xt)


def test_rk4_cart_and_mecanum_bit_identical_to_reference():
    dt, steps = float(G["dt"]), int(G["steps"])
    cp, mp = G["cart_params"], G["mecanum_params"]
    for i in range(G["x0"].shape[0]):
        a = Oracle.rk4_forward_cart(*cp, dt, steps * dt, G["x0"][i], G["ut_cart"][i])
        assert np.array_equal(a, G["xt_cart"][i])
        b = Oracle.rk4_forward_mecanum(*mp, dt, steps * dt, G["x0"][i], G["ut_mecanum"][i])
        assert np.array_equal(b, G["xt_mecanum"][i])


def test_model_jacobians_bit_identical_to_reference():
    cp, mp = G["cart_params"], G["mecanum_params"]
    for i in range(G["x0"].shape[0]):
        f, A, B, vb = Oracle.cart(*cp, G["x0"][i], G["ut_cart"][i, 0])
        for got, key in ((f, "f"), (A, "A"), (B, "B"), (vb, "vb")):
            assert np.array_equal(got, G["cart_" + key][i]), key
        f, A, B, vb = Oracle.mecanum(*mp, G["x0"][i], G["ut_mecanum"][i, 0])
        for got, key in ((f, "f"), (A, "A"), (B, "B"), (vb, "vb")):
            assert np.array_equal(got, G["mecanum_" + key][i]), key


def test_entropy_grid_bit_identical_to_reference():
    e = Oracle.entropy_grid(G["cells"])
    assert np.array_equal(e, G["entropy"])
    # numerics.hpp:164-179: p = 0 and p = 1 -> 1e-3, unknown (-1 -> p < 0) -> 0.7, p = 0.5 -> ln 2
    assert e[0, 0] == 1e-3 and e[0, 1] == 1e-3 and e[0, 2] == 0.7
    assert abs(e[0, 3] - np.log(2.0)) < 1e-15
