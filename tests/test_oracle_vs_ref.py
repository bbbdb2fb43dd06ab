"""The plain-C restatement against the UNMODIFIED reference sources compiled
against the test shim (oracle/_ref).  This is what pins the oracle for the
parts the reference's own tests do not cover (Basis, Target, ReplayBuffer,
ErgodicControl, backward RK4, Omni).  Skipped where oracle/_ref is absent."""
import numpy as np
import pytest

from helpers import BOUNDS_10, BOUNDS_MAZE, MODEL_OMNI, MODEL_SIMPLE_CART, make_oracle, plant, random_states, warm_ut
from oracle.pyoracle import Oracle, RefLib

pytestmark = pytest.mark.skipif(not RefLib.available(), reason="compiled reference oracle/_ref not built")

TIGHT = 1e-12  # the restatement follows the reference's association order: expect ~0


def test_models_and_rk4():
    rng = np.random.default_rng(1)
    for model in (MODEL_SIMPLE_CART, MODEL_OMNI):
        for _ in range(20):
            x = rng.uniform(-3, 3, 3)
            u = rng.uniform(-2, 2, 3)
            if model == MODEL_SIMPLE_CART:
                u[1] = 0.0
            np.testing.assert_allclose(Oracle.model_f(model, x, u), RefLib.model_f(model, x, u), rtol=0, atol=TIGHT)
            np.testing.assert_allclose(Oracle.model_fdx(model, x, u), RefLib.model_fdx(model, x, u), rtol=0, atol=TIGHT)
            np.testing.assert_allclose(Oracle.model_fdu(model, x), RefLib.model_fdu(model, x), rtol=0, atol=TIGHT)
        ut = rng.uniform(-1, 1, (37, 3))
        if model == MODEL_SIMPLE_CART:
            ut[:, 1] = 0.0
        a = Oracle.rk4_forward(model, 0.1, 3.7, [0.5, -0.2, 3.0], ut)
        b = RefLib.rk4_forward(model, 0.1, 3.7, [0.5, -0.2, 3.0], ut)
        np.testing.assert_allclose(a, b, rtol=0, atol=TIGHT)


@pytest.mark.parametrize("nb", [1, 3, 10, 17])
def test_basis(nb):
    rng = np.random.default_rng(nb)
    lx, ly = 7.3, 12.9
    for _ in range(5):
        x = rng.uniform(0, 7, 2)
        np.testing.assert_allclose(Oracle.fourier_basis(lx, ly, nb, x), RefLib.fourier_basis(lx, ly, nb, x), rtol=0, atol=TIGHT)
        np.testing.assert_allclose(Oracle.grad_fourier_basis(lx, ly, nb, x), RefLib.grad_fourier_basis(lx, ly, nb, x), rtol=0, atol=TIGHT)
    xt = rng.uniform(0, 7, (23, 3))
    np.testing.assert_allclose(Oracle.traj_coeff(lx, ly, nb, xt), RefLib.traj_coeff(lx, ly, nb, xt), rtol=0, atol=TIGHT)
    ka, la = Oracle.basis_tables(nb)
    kb, lb = RefLib.basis_tables(nb)
    assert (ka == kb).all() and (la == lb).all()


def test_target_fill_and_spatial_coeff():
    grid, nx, ny = Oracle.target_grid(6.0, 4.3, 0.1)
    assert (nx, ny) == (61, 44)
    mu, sg, trans = [[2.5, 2.5], [4.0, 1.0], [5.5, 3.9]], [[1.5, 1.5], [0.3, 0.9], [0.7, 0.2]], [-1.0, 0.5]
    va = Oracle.target_fill(mu, sg, trans, grid)
    vb = RefLib.target_fill(mu, sg, trans, grid)
    np.testing.assert_allclose(va, vb, rtol=1e-14, atol=0)
    assert abs(va.sum() - 1.0) < 1e-12
    pa = Oracle.spatial_coeff(6.0, 4.3, 8, va, grid)
    pb = RefLib.spatial_coeff(6.0, 4.3, 8, vb, grid)
    np.testing.assert_allclose(pa, pb, rtol=0, atol=TIGHT)
    # the streaming restatement used for big grids is the same arithmetic
    pc, _ = Oracle.phik_from_grid(va.reshape(ny, nx) * 3.7, 0.1, 6.0, 4.3, 8)
    np.testing.assert_allclose(pc, pa, rtol=0, atol=1e-13)


@pytest.mark.parametrize("model", [MODEL_SIMPLE_CART, MODEL_OMNI])
@pytest.mark.parametrize("bounds", [BOUNDS_10, BOUNDS_MAZE])
def test_control_closed_loop(model, bounds):
    """config C1: 8 closed-loop control() steps with growing replay memory;
    control() itself and the step-by-step trace of the same member calls"""
    mu = np.array([[2.5, 2.5], [8.5, 2.5]]) + np.array([bounds[0], bounds[2]])
    o = make_oracle(model, mu=mu)
    r = make_oracle(model, mu=mu, lib=RefLib)
    r2 = make_oracle(model, mu=mu, lib=RefLib)
    x = np.array([bounds[0] + 5.0, bounds[2] + 7.0, 0.3])
    for step in range(8):
        uo = o.control(bounds, x)
        ur = r.control(bounds, x)
        ut = r2.control(bounds, x, trace=True)
        np.testing.assert_allclose(uo, ur, rtol=0, atol=TIGHT)
        np.testing.assert_array_equal(ur, ut)
        lo, lr = o.last(), r2.last()
        for key in ("ck", "edx", "bdx", "rhot", "xtf"):
            np.testing.assert_allclose(lo[key], lr[key], rtol=0, atol=TIGHT, err_msg=key)
        np.testing.assert_allclose(o.get_ut(), r.get_ut(), rtol=0, atol=TIGHT)
        np.testing.assert_allclose(o.opt_traj(), r.opt_traj(), rtol=0, atol=TIGHT)
        x = plant(x[None], uo[None])[0]
        for c in (o, r, r2):
            c.add_state_memory(x)
    np.testing.assert_allclose(o.get_phik(), r.get_phik(), rtol=0, atol=TIGHT)


def test_control_with_full_replay_batch():
    """memory == batch_size: the deterministic 'all stored states' branch (buffer.cpp:75-89)"""
    rng = np.random.default_rng(4)
    o = make_oracle(MODEL_OMNI, batch_size=100)
    r = make_oracle(MODEL_OMNI, batch_size=100, lib=RefLib)
    for s in random_states(rng, 100):
        o.add_state_memory(s)
        r.add_state_memory(s)
    ut = warm_ut(rng, 1, 50, MODEL_OMNI)[0]
    o.set_ut(ut)
    r.set_ut(ut)
    x = np.array([3.0, 4.0, -1.0])
    np.testing.assert_allclose(o.control(BOUNDS_10, x), r.control(BOUNDS_10, x), rtol=0, atol=TIGHT)


def test_explicit_indices_equal_prefilled_buffer():
    """feeding sample indices to the oracle == a reference
    buffer pre-filled with exactly those states (batch_size >= count)"""
    rng = np.random.default_rng(8)
    hist = random_states(rng, 300)
    idx = rng.integers(0, 300, 100)
    o = make_oracle(MODEL_OMNI, batch_size=100)
    for s in hist:
        o.add_state_memory(s)
    r = make_oracle(MODEL_OMNI, batch_size=100, lib=RefLib)
    for s in hist[idx]:
        r.add_state_memory(s)
    x = np.array([6.0, 2.0, 0.5])
    np.testing.assert_allclose(o.control(BOUNDS_10, x, mem_idx=idx), r.control(BOUNDS_10, x), rtol=0, atol=TIGHT)
