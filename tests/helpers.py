"""Shared test helpers: reference configurations, seeded
inputs, and the parity tolerances.

Tolerances (BASELINE.json north_star: 1e-9 relative in FP64):
  * coefficient vectors (phi_k, c_k):  max|a-b| / max|b| <= 1e-9
    (phi_k spans > 5 decades, element-wise relative error is meaningless)
  * controls, states, co-states, metric:  |a-b| <= 1e-9 * max(1, |b|)
"""
import numpy as np

from oracle.pyoracle import MODEL_OMNI, MODEL_SIMPLE_CART, Oracle

RTOL = 1e-9

# config/explore_{cart,omni}.yaml + node mains
TARGET_MU = np.array([[2.5, 2.5], [8.5, 2.5]])
TARGET_SIGMA = np.array([[1.5, 1.5], [1.5, 1.5]])
BOUNDS_10 = (0.0, 10.0, 0.0, 10.0)
# maps/maze.yaml: origin (-8.350015, -13.950030), 199 x 318 cells @ 0.1
BOUNDS_MAZE = (-8.350015, -8.350015 + 19.9, -13.950030, -13.950030 + 31.8)


def model_params(model):
    if model == MODEL_OMNI:  # explore_omni.yaml:18-24,56; exploration_omni_node.cpp:161-164
        return np.diag([1.0, 1.0, 2.0]), np.array([-1.0, -1.0, -2.0]), np.array([1.0, 1.0, 2.0])
    # explore_cart.yaml:15-19,50; exploration_cart_node.cpp:134-135,156-159
    return np.diag([1.0, 0.0, 2.0]), np.array([-1.0, 0.0, -2.0]), np.array([1.0, 0.0, 2.0])


def make_oracle(model, nb=10, horizon=5.0, dt=0.1, batch_size=100, buffer_size=1000000, res=0.1, w=1.0,
                lib=Oracle, mu=TARGET_MU, sigma=TARGET_SIGMA):
    R, umin, umax = model_params(model)
    c = lib.create(model, dt, horizon, res, w, nb, buffer_size, batch_size, R, umin, umax)
    c.set_target(mu, sigma)
    return c


def make_gpu(model, batch, nb=10, horizon=5.0, dt=0.1, batch_size=100, buffer_size=1000000, res=0.1, w=1.0,
             mu=TARGET_MU, sigma=TARGET_SIGMA, **kw):
    import ergodic_exploration_b200 as eb

    R, umin, umax = model_params(model)
    c = eb.ErgodicControl(model, dt, horizon, res, w, nb, buffer_size, batch_size, R, umin, umax, batch=batch, **kw)
    c.setTarget([eb.Gaussian(m, s) for m, s in zip(mu, sigma)])
    return c


def random_states(rng, n, bounds=BOUNDS_10, margin=0.5):
    xmin, xmax, ymin, ymax = bounds
    x = np.empty((n, 3))
    x[:, 0] = rng.uniform(xmin + margin, xmax - margin, n)
    x[:, 1] = rng.uniform(ymin + margin, ymax - margin, n)
    x[:, 2] = rng.uniform(-np.pi, np.pi, n)
    return x


def warm_ut(rng, n, steps, model):
    """ut_ warm state: each component ~ U(umin, umax) * 0.5"""
    _, umin, umax = model_params(model)
    return rng.uniform(umin, umax, size=(n, steps, 3)) * 0.5


def plant(x, u, dt=0.1):
    """the reference's constant-twist integrator + angle wrap (numerics.hpp:273-298,77-89)"""
    out = np.empty_like(x)
    for i in range(x.shape[0]):
        xn = Oracle.integrate_twist(x[i], u[i], dt)
        xn[2] = Oracle.normalize_angle_pi(xn[2])
        out[i] = xn
    return out


def assert_coeff_close(a, b, what="", rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.max(np.abs(b))
    err = np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0)
    assert err <= rtol, f"{what}: max|a-b|/max|b| = {err:.3e} > {rtol:g}"


def assert_abs_rel_close(a, b, what="", rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert np.all(np.isfinite(a)), f"{what}: non-finite values"
    assert err.max() <= rtol, f"{what}: max |a-b|/max(1,|b|) = {err.max():.3e} > {rtol:g}"


def assert_angle_close(a, b, what="", rtol=RTOL):
    """headings compared modulo 2*pi (the wrap boundary -pi / pi is one point)"""
    d = np.asarray(a) - np.asarray(b)
    d = np.abs(np.arctan2(np.sin(d), np.cos(d)))
    assert d.max() <= 10 * rtol, f"{what}: max angular error {d.max():.3e}"


def oracle_phik_threaded(phi, res, lx, ly, nb, threads=None):
    """Oracle.phik_from_grid's arithmetic (the restated spatialCoeff, every cell x every basis function) with the rows
    spread over the host cores: eo_phik_rows per row block (ctypes releases the GIL), partials summed in block order.
    For the large-grid / wide-basis cases, where the single-threaded call takes minutes."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    phi = np.ascontiguousarray(phi, dtype=np.float64)
    ny = phi.shape[0]
    total = float(phi.sum())
    threads = threads or min(32, os.cpu_count() or 1)
    bounds = np.linspace(0, ny, threads + 1).astype(int)
    blocks = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    with ThreadPoolExecutor(len(blocks)) as ex:
        parts = list(ex.map(lambda ab: Oracle.phik_rows(phi[ab[0]:ab[1]], ab[0], res, lx, ly, nb, total), blocks))
    return np.sum(parts, axis=0), total
