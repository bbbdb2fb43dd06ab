"""-m gpu: the kinematic models and the forward RK4 integrator on the GPU (csrc/model_kernels.cuh) against the golden
vectors of the compiled reference and the CPU oracle -- SURVEY.md section 8 rows a8 / a9.  The reference's own tests:
test/test_integrator.cpp:44-73 (Cart RK4), test/test_cart.cpp:42-167 and test/test_omni.cpp:42-104 (f, A, B at
x = {1, 2, 0.707}).  Tolerance 1e-9 (only the last ulp of sin / cos differs from the host libm)."""
import os

import numpy as np
import pytest

from helpers import assert_abs_rel_close, assert_angle_close
from oracle.pyoracle import MODEL_OMNI, MODEL_SIMPLE_CART, Oracle

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_entropy.npz"))


def test_reference_integrator_test_on_gpu():
    """test/test_integrator.cpp:44-73"""
    import ergodic_exploration_b200 as eb

    xt = eb.RungeKutta(0.1).solve(eb.Cart(0.1, 2.0), [0.0, 0.0, 0.0], np.ones((4, 2)), 0.4)
    assert xt.shape == (4, 3)
    for i in range(4):
        assert abs(xt[i, 0] - 0.01 * (i + 1)) <= 4 * np.spacing(0.01 * (i + 1))
        assert xt[i, 1] == 0.0 and xt[i, 2] == 0.0


def test_reference_model_known_answers_on_gpu():
    """test/test_cart.cpp:132-167 (SimpleCart), :42-88 (Cart), test/test_omni.cpp:42-104 (Mecanum), 1e-6 as there"""
    import ergodic_exploration_b200 as eb

    x = [1.0, 2.0, 0.707]
    sc = eb.SimpleCart()
    u = [0.5, 0.0, 0.01]
    assert np.allclose(sc(x, u), [0.380156, 0.324777, 0.01], atol=1e-6)
    A = sc.fdx(x, u)
    assert abs(A[0, 2] + 0.324777) < 1e-6 and abs(A[1, 2] - 0.380156) < 1e-6
    B = sc.fdu(x)
    assert abs(B[0, 0] - 0.760313) < 1e-6 and abs(B[1, 0] - 0.649555) < 1e-6 and B[2, 2] == 1.0
    with pytest.raises(ValueError):
        sc(x, [0.5, 0.1, 0.01])  # cart.hpp:167-170
    # Cart / Mecanum against the CPU restatement at the same point (pinned to the reference's KATs in test_oracle_kat.py)
    for model, ref, uu in ((eb.Cart(0.033, 0.08), lambda: Oracle.cart(0.033, 0.08, x, [1.0, 0.5]), [1.0, 0.5]),
                           (eb.Mecanum(0.1, 0.5, 0.3), lambda: Oracle.mecanum(0.1, 0.5, 0.3, x, [1.0, 0.5, 0.3, 0.2]),
                            [1.0, 0.5, 0.3, 0.2])):
        f, A, B, vb = ref()
        assert_abs_rel_close(model(x, uu), f, "f")
        assert_abs_rel_close(model.fdx(x, uu), A, "fdx")
        assert_abs_rel_close(model.fdu(x), B, "fdu")
        assert_abs_rel_close(model.wheels2Twist(uu), vb, "wheels2Twist")


def test_rk4_cart_mecanum_against_reference_golden():
    import ergodic_exploration_b200 as eb

    dt, steps = float(G["dt"]), int(G["steps"])
    rk = eb.RungeKutta(dt)
    for model, ut, want in ((eb.Cart(*G["cart_params"]), G["ut_cart"], G["xt_cart"]),
                            (eb.Mecanum(*G["mecanum_params"]), G["ut_mecanum"], G["xt_mecanum"])):
        got = rk.solve(model, G["x0"], ut, steps * dt)  # batched, one control signal per instance
        assert got.shape == want.shape
        assert_abs_rel_close(got[:, :, :2], want[:, :, :2], "positions")
        assert_angle_close(got[:, :, 2], want[:, :, 2], "headings")
        one = rk.solve(model, G["x0"][3], ut[3], steps * dt)  # single instance, Armadillo-style call
        assert np.array_equal(one, got[3])


def test_model_batch_against_reference_golden():
    import ergodic_exploration_b200 as eb

    for model, name, ut in ((eb.Cart(*G["cart_params"]), "cart", G["ut_cart"]),
                            (eb.Mecanum(*G["mecanum_params"]), "mecanum", G["ut_mecanum"])):
        u0 = np.ascontiguousarray(ut[:, 0])
        assert_abs_rel_close(model(G["x0"], u0), G[name + "_f"], name + " f")
        assert_abs_rel_close(model.fdx(G["x0"], u0), G[name + "_A"], name + " A")
        assert_abs_rel_close(model.fdu(G["x0"]), G[name + "_B"], name + " B")
        assert_abs_rel_close(model.wheels2Twist(u0), G[name + "_vb"], name + " vb")


@pytest.mark.parametrize("model_id", [MODEL_SIMPLE_CART, MODEL_OMNI])
def test_sequential_rk4_agrees_with_the_scan_rollout(model_id):
    """the one-thread-per-instance integrator and the warp-scan rollout of the fused kernel (optTraj) are two
    independent formulations of integrator.hpp:135-152: both must match the oracle"""
    import ergodic_exploration_b200 as eb

    rng = np.random.default_rng(5 + model_id)
    n, steps, dt = 40, 70, 0.1
    x0 = np.column_stack([rng.uniform(1, 9, n), rng.uniform(1, 9, n), rng.uniform(-np.pi, np.pi, n)])
    ut = rng.uniform(-1, 1, (n, steps, 3))
    if model_id == MODEL_SIMPLE_CART:
        ut[:, :, 1] = 0.0
    model = eb.SimpleCart() if model_id == MODEL_SIMPLE_CART else eb.Omni()
    got = eb.RungeKutta(dt).solve(model, x0, ut, steps * dt)
    for i in range(0, n, 7):
        want = Oracle.rk4_forward(model_id, dt, steps * dt, x0[i], ut[i])
        assert_abs_rel_close(got[i, :, :2], want[:, :2], "positions")
        assert_angle_close(got[i, :, 2], want[:, 2], "headings")
