"""The CPU oracle against the reference's OWN known-answer vectors
(SURVEY.md §8c): test/test_cart.cpp, test/test_omni.cpp, test/test_integrator.cpp.
Tolerances are the reference tests' own (ASSERT_NEAR 1e-6 / ASSERT_DOUBLE_EQ)."""
import numpy as np
import pytest

from oracle.pyoracle import MODEL_OMNI, MODEL_SIMPLE_CART, Oracle, RefLib

LIBS = [Oracle] + ([RefLib] if RefLib.available() else [])


@pytest.mark.parametrize("lib", LIBS, ids=lambda l: l.__name__)
def test_cart_kinematics(lib):
    """test/test_cart.cpp:42-130 (Cart(0.1, 2.0) at x = {1, 2, 0.707})"""
    x = [1.0, 2.0, 0.707]
    f, A, B, _ = lib.cart(0.1, 2.0, x, [1.0, 0.5])
    # kinematics / fdx / fdu recomputed from the closed forms the test literals come from
    r, c, s = 0.1, np.cos(0.707), np.sin(0.707)
    np.testing.assert_allclose(f, [r / 2 * 1.5 * c, r / 2 * 1.5 * s, r / 2 * (0.5 - 1.0) / 2.0], atol=1e-6)
    np.testing.assert_allclose(A[:, 2], [-r / 2 * 1.5 * s, r / 2 * 1.5 * c, 0.0], atol=1e-6)
    np.testing.assert_allclose(B, [[r / 2 * c] * 2, [r / 2 * s] * 2, [-r / 4, r / 4]], atol=1e-6)
    # wheels2Twist: forward, rotate in place, arc (test_cart.cpp:90-130)
    for u, want in (([1.0, 1.0], [0.1, 0.0, 0.0]), ([-1.0, 1.0], [0.0, 0.0, 0.05]),
                    ([0.0, 1.0], [0.05, 0.0, 0.025])):
        np.testing.assert_allclose(lib.cart(0.1, 2.0, x, u)[3], want, atol=1e-6)


@pytest.mark.parametrize("lib", LIBS, ids=lambda l: l.__name__)
def test_simple_cart_literals(lib):
    """test/test_cart.cpp:132-167: the 6-digit literals quoted in SURVEY §8c"""
    x, u = [1.0, 2.0, 0.707], [0.5, 0.0, 0.01]
    np.testing.assert_allclose(lib.model_f(MODEL_SIMPLE_CART, x, u), [0.380156, 0.324777, 0.01], atol=1e-6)
    A = lib.model_fdx(MODEL_SIMPLE_CART, x, u)
    assert abs(A[0, 2] - (-0.324777)) < 1e-6 and abs(A[1, 2] - 0.380156) < 1e-6
    assert np.count_nonzero(A) == 2
    B = lib.model_fdu(MODEL_SIMPLE_CART, x)
    assert abs(B[0, 0] - 0.760313) < 1e-6 and abs(B[1, 0] - 0.649555) < 1e-6 and B[2, 2] == 1.0
    assert np.count_nonzero(B) == 3


@pytest.mark.parametrize("lib", LIBS, ids=lambda l: l.__name__)
def test_simple_cart_rejects_lateral_velocity(lib):
    """cart.hpp:167-170"""
    with pytest.raises(ValueError):
        lib.model_f(MODEL_SIMPLE_CART, [0, 0, 0], [0.5, 0.1, 0.0])


@pytest.mark.parametrize("lib", LIBS, ids=lambda l: l.__name__)
def test_mecanum_kinematics(lib):
    """test/test_omni.cpp:42-104 (Mecanum(0.1, 0.5, 0.5))"""
    x, u = [1.0, 2.0, 0.707], [1.0, 0.5, 0.25, 0.75]
    f, A, B, tw = lib.mecanum(0.1, 0.5, 0.5, x, u)
    s, c, l = 0.025 * np.sin(0.707), 0.025 * np.cos(0.707), 0.1 / 4.0
    Bw = np.array([[s + c, -s + c, s + c, -s + c], [s - c, s + c, s - c, s + c], [-l, l, l, -l]])
    np.testing.assert_allclose(B, Bw, atol=1e-6)
    np.testing.assert_allclose(f, Bw @ np.array(u), atol=1e-6)
    dB = np.array([[c - s, -c - s, c - s, -c - s], [c + s, c - s, c + s, c - s]])
    np.testing.assert_allclose(A[:2, 2], dB @ np.array(u), atol=1e-6)
    np.testing.assert_allclose(tw, 0.025 * np.array([[1, 1, 1, 1], [-1, 1, -1, 1], [-1, 1, 1, -1]]) @ np.array(u),
                               atol=1e-6)


@pytest.mark.parametrize("lib", LIBS, ids=lambda l: l.__name__)
def test_rk4_straight_line(lib):
    """test/test_integrator.cpp:44-73: ASSERT_DOUBLE_EQ (4 ulp)"""
    xt = lib.rk4_forward_cart(0.1, 2.0, 0.1, 0.4, [0.0, 0.0, 0.0], np.ones((4, 2)))
    for i in range(4):
        want = 0.01 * (i + 1)
        assert abs(xt[i, 0] - want) <= 4 * np.spacing(want)
        assert xt[i, 1] == 0.0 and xt[i, 2] == 0.0


@pytest.mark.parametrize("lib", LIBS, ids=lambda l: l.__name__)
def test_integrate_twist_pose_change(lib):
    """test/test_integrator.cpp:75-142: constant twist vs closed form (1e-4 there)"""
    x = np.zeros(3)
    for _ in range(10):
        x = lib.integrate_twist(x, [1.0, 0.5, 0.5], 0.1)
    w, t = 0.5, 1.0
    want = [(1.0 * np.sin(w * t) + 0.5 * (np.cos(w * t) - 1.0)) / w,
            (0.5 * np.sin(w * t) + 1.0 * (1.0 - np.cos(w * t))) / w, w * t]
    np.testing.assert_allclose(x, want, atol=1e-4)
    np.testing.assert_allclose(lib.integrate_twist([1.0, 2.0, 0.3], [0.4, -0.2, 0.0], 0.1),
                               [1.0 + 0.1 * (0.4 * np.cos(0.3) + 0.2 * np.sin(0.3)),
                                2.0 + 0.1 * (0.4 * np.sin(0.3) - 0.2 * np.cos(0.3)), 0.3], atol=1e-12)


@pytest.mark.parametrize("lib", LIBS, ids=lambda l: l.__name__)
def test_normalize_angle(lib):
    for r, want in ((0.0, 0.0), (np.pi / 2, np.pi / 2), (3 * np.pi / 2, -np.pi / 2), (-3 * np.pi / 2, np.pi / 2),
                    (5 * np.pi, -np.pi), (-7.5 * np.pi, np.pi / 2)):
        assert abs(lib.normalize_angle_pi(r) - want) < 1e-12


@pytest.mark.parametrize("lib", LIBS, ids=lambda l: l.__name__)
def test_basis_tables(lib):
    """basis.cpp:48-77: index = ky*nb + kx, lamda = 1/(1+|k|)^1.5"""
    k, lam = lib.basis_tables(4)
    assert k[:, 0].tolist() == [0, 1, 2, 3] * 4 and k[:, 1].tolist() == sum(([i] * 4 for i in range(4)), [])
    np.testing.assert_allclose(lam, 1.0 / (1.0 + np.sqrt((k ** 2).sum(1))) ** 1.5, rtol=1e-15)
    assert lam[0] == 1.0


def test_ctor_needs_two_steps():
    """ergodic_control.hpp:212-216"""
    with pytest.raises(ValueError):
        Oracle.create(MODEL_OMNI, 0.1, 0.1, 0.1, 1.0, 4, 10, 10, np.eye(3), [-1] * 3, [1] * 3)
    c = Oracle.create(MODEL_OMNI, 0.1, 0.3, 0.1, 1.0, 4, 10, 10, np.eye(3), [-1] * 3, [1] * 3)
    assert c.steps == 2  # (unsigned)|0.3/0.1| truncates
