"""CPU: the collision restatement (oracle/ergodic_oracle.c) against the compiled
reference (Collision, GridMap, validate_control built from the unmodified sources),
and the geometric fact the CUDA kernel's radius pruning relies on."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle, RefLib

needs_ref = pytest.mark.skipif(not RefLib.available(), reason="compiled reference (oracle/_ref) not present")


def random_map(rng, ys, xs, p_occ=0.01, p_unknown=0.02):
    """int8 occupancy: 0 free, 100 occupied, -1 unknown, a few mid-probability cells and a wall"""
    data = np.zeros((ys, xs), dtype=np.int8)
    data[rng.random((ys, xs)) < p_occ] = 100
    data[rng.random((ys, xs)) < p_unknown] = -1
    data[rng.random((ys, xs)) < 0.01] = rng.integers(1, 100)
    data[ys // 3: ys // 3 + 3, xs // 5: 4 * xs // 5] = 100
    return data


def reference_circle(r):
    """cells (dx, dy) visited by Collision::bresenhamCircle (collision.cpp:169-215)"""
    x, y, err = -r, 0, 2 - 2 * r
    pts = []
    while x < 0:
        pts += [(-x, y), (-y, -x), (x, -y), (y, x)]
        rr = err
        if rr <= y:
            y += 1
            err += 2 * y + 1
        if rr > x or err > y:
            x += 1
            err += 2 * x + 1
    return pts


def test_circle_cells_lie_outside_radius_minus_one():
    """every cell of the radius-r walk is farther than r - 1 from the centre, so a walk
    with r >= r_col + 1 cannot contain a cell with dx^2 + dy^2 <= r_col^2: the kernel may stop
    the search at min(r_max, r_col).  Checked for every radius the kernel prunes (< 3000)."""
    for r in range(1, 3000):
        assert min(a * a + b * b for a, b in reference_circle(r)) > (r - 1) ** 2, r


CASES = [
    # (ysize, xsize, res, xmin, ymin, (boundary, search, obstacle_thr, occupied_thr))
    (120, 160, 0.05, -1.0, 2.0, (0.2, 1.0, 0.05, 0.9)),     # explore_*.yaml-like radii
    (64, 48, 0.1, 0.0, 0.0, (0.25, 0.25, 0.0, 0.5)),        # search == boundary, one circle
    (200, 200, 0.05, -5.0, -5.0, (0.03, 0.4, 0.3, 0.65)),   # r_bnd = 0, r_col > r_bnd
    (90, 130, 0.2, 3.3, -7.1, (0.5, 2.0, 0.1, 0.0)),        # threshold 0: every non-negative cell counts
    (70, 70, 0.1, 0.0, 0.0, (0.3, 0.6, -0.55, 0.5)),        # negative collision radius: r_col^2 still compares
]


@needs_ref
@pytest.mark.parametrize("case", CASES)
def test_collision_check_restatement_matches_reference(case):
    ys, xs, res, xmin, ymin, col = case
    rng = np.random.default_rng(ys * xs)
    data = random_map(rng, ys, xs)
    n = 3000
    poses = np.column_stack([rng.uniform(xmin - 0.7, xmin + xs * res + 0.7, n),
                             rng.uniform(ymin - 0.7, ymin + ys * res + 0.7, n), rng.uniform(-3.1, 3.1, n)])
    poses[0, :2] = (xmin + xs * res, ymin + ys * res)  # exactly on the upper edge (grid.cpp:148-156)
    poses[1, :2] = (xmin, ymin)
    a = Oracle.collision_check(data, res, xmin, ymin, col, poses)
    b = RefLib.collision_check(data, res, xmin, ymin, col, poses)
    np.testing.assert_array_equal(a, b)
    assert 0 < a.sum() < n


@needs_ref
@pytest.mark.parametrize("case", CASES)
def test_validate_control_restatement_matches_reference(case):
    ys, xs, res, xmin, ymin, col = case
    rng = np.random.default_rng(7 * ys + xs)
    data = random_map(rng, ys, xs)
    n = 1500
    x0 = np.column_stack([rng.uniform(xmin, xmin + xs * res, n), rng.uniform(ymin, ymin + ys * res, n),
                          rng.uniform(-np.pi, np.pi, n)])
    u = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-2, 2, n)])
    u[::5, 2] = 0.0  # the no-rotation branch of integrate_twist (numerics.hpp:280-285)
    for dt, horizon in ((0.1, 0.5), (0.1, 2.0), (0.05, 0.33)):
        a = Oracle.validate_control(data, res, xmin, ymin, col, x0, u, dt, horizon)
        b = RefLib.validate_control(data, res, xmin, ymin, col, x0, u, dt, horizon)
        np.testing.assert_array_equal(a, b)


@needs_ref
def test_reference_constructor_errors():
    data = np.zeros((4, 4), dtype=np.int8)
    with pytest.raises(ValueError):  # search radius < boundary radius (collision.cpp:52-56)
        RefLib.collision_check(data, 0.1, 0.0, 0.0, (0.5, 0.2, 0.0, 0.5), np.zeros((1, 3)))
    with pytest.raises(ValueError):  # occupied threshold outside [0, 100] (collision.cpp:58-61)
        RefLib.collision_check(data, 0.1, 0.0, 0.0, (0.1, 0.2, 0.0, 101.0), np.zeros((1, 3)))
