"""-m gpu: batched Collision::collisionCheck / validate_control on the GPU (through the
C ABI) against the CPU restatement.  Integer / index work: the results must be
identical (int arrays compared with assert_array_equal)."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle
from test_collision_cpu import CASES, random_map

pytestmark = pytest.mark.gpu


def _grid(data, res, xmin, ymin):
    import ergodic_exploration_b200 as eb

    ys, xs = data.shape
    return eb.GridMap(xmin, xmin + xs * res, ymin, ymin + ys * res, res, data)


@pytest.mark.parametrize("case", CASES)
def test_collision_check_matches_cpu(case):
    import ergodic_exploration_b200 as eb

    ys, xs, res, xmin, ymin, col = case
    rng = np.random.default_rng(ys + 31 * xs)
    data = random_map(rng, ys, xs)
    grid, c = _grid(data, res, xmin, ymin), eb.Collision(*col)
    n = 20000
    poses = np.column_stack([rng.uniform(xmin - 0.7, xmin + xs * res + 0.7, n),
                             rng.uniform(ymin - 0.7, ymin + ys * res + 0.7, n), rng.uniform(-3.1, 3.1, n)])
    poses[0, :2] = (xmin + xs * res, ymin + ys * res)
    poses[1, :2] = (xmin, ymin)
    want = Oracle.collision_check(data, res, xmin, ymin, col, poses)
    for mode in (1, 2, 0):  # circle walk, pre-dilated map, automatic
        grid.dilation(mode)
        np.testing.assert_array_equal(c.collisionCheck(grid, poses), want)
    assert 0 < want.sum() < n


@pytest.mark.parametrize("case", CASES)
def test_validate_control_matches_cpu(case):
    import ergodic_exploration_b200 as eb

    ys, xs, res, xmin, ymin, col = case
    rng = np.random.default_rng(5 * ys + xs)
    data = random_map(rng, ys, xs)
    grid, c = _grid(data, res, xmin, ymin), eb.Collision(*col)
    n = 8000
    x0 = np.column_stack([rng.uniform(xmin, xmin + xs * res, n), rng.uniform(ymin, ymin + ys * res, n),
                          rng.uniform(-np.pi, np.pi, n)])
    u = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-2, 2, n)])
    u[::5, 2] = 0.0
    for dt, horizon in ((0.1, 0.5), (0.1, 2.0), (0.05, 0.33)):
        want = Oracle.validate_control(data, res, xmin, ymin, col, x0, u, dt, horizon)
        for mode in (1, 2):
            grid.dilation(mode)
            np.testing.assert_array_equal(eb.validate_control(c, grid, x0, u, dt, horizon), want)
        assert want.sum() < n and (want.sum() > 0 or col[3] == 0.0)  # threshold 0: every known cell is an obstacle


def test_device_path_update_and_edge_cases():
    """torch tensors in/out on the current stream; GridMap::update; empty map -> nothing collides;
    full map -> everything inside the search disc collides; zero-length batch"""
    import torch

    import ergodic_exploration_b200 as eb

    rng = np.random.default_rng(3)
    ys, xs, res = 80, 100, 0.1
    col = (0.2, 0.6, 0.1, 0.5)
    free = np.zeros((ys, xs), dtype=np.int8)
    grid, c = _grid(free, res, 0.0, 0.0), eb.Collision(*col)
    poses = np.column_stack([rng.uniform(0, xs * res, 4096), rng.uniform(0, ys * res, 4096), np.zeros(4096)])
    pd = torch.from_numpy(poses).cuda()
    assert int(c.collisionCheck(grid, pd).sum()) == 0
    full = np.full((ys, xs), 100, dtype=np.int8)
    grid.update(full)
    hit = c.collisionCheck(grid, pd).cpu().numpy()
    np.testing.assert_array_equal(hit, Oracle.collision_check(full, res, 0.0, 0.0, col, poses))
    assert hit.all()
    u = torch.from_numpy(np.tile([0.3, 0.0, 0.2], (4096, 1))).cuda()
    v = eb.validate_control(c, grid, pd, u, 0.1, 0.5).cpu().numpy()
    assert not v.any()
    assert len(c.collisionCheck(grid, np.zeros((0, 3)))) == 0
    assert grid.launch_count() >= 3


def test_collision_constructor_and_grid_errors():
    import ergodic_exploration_b200 as eb

    with pytest.raises(ValueError):
        eb.Collision(0.5, 0.2, 0.0, 0.5)
    with pytest.raises(ValueError):
        eb.Collision(0.1, 0.2, 0.0, 101.0)
    with pytest.raises(ValueError):  # grid.cpp:57-60
        eb.GridMap(0.0, 1.0, 0.0, 1.0, 0.1, np.zeros(99, dtype=np.int8))


def test_validate_control_large_map_properties():
    """BASELINE-scale batch on a 4000 x 4000 map (no CPU run at this size): a twist that is
    valid over a horizon is valid over every shorter horizon, and zero twist == collisionCheck
    of the start pose"""
    import torch

    import ergodic_exploration_b200 as eb

    rng = np.random.default_rng(11)
    n, res = 4000, 0.05
    data = np.zeros((n, n), dtype=np.int8)
    data[rng.random((n, n)) < 0.002] = 100
    grid, c = _grid(data, res, -100.0, -100.0), eb.Collision(0.2, 1.0, 0.05, 0.65)
    B = 1 << 18
    x0 = torch.from_numpy(np.column_stack([rng.uniform(-99, 99, B), rng.uniform(-99, 99, B),
                                           rng.uniform(-np.pi, np.pi, B)])).cuda()
    u = torch.from_numpy(np.column_stack([rng.uniform(-1, 1, B), rng.uniform(-1, 1, B), rng.uniform(-2, 2, B)])).cuda()
    long_ok = eb.validate_control(c, grid, x0, u, 0.1, 2.0)
    short_ok = eb.validate_control(c, grid, x0, u, 0.1, 0.5)
    assert bool((short_ok >= long_ok).all()) and 0 < int(long_ok.sum()) < B
    zero = torch.zeros_like(u)
    still = eb.validate_control(c, grid, x0, zero, 0.1, 0.1)
    hit = c.collisionCheck(grid, x0)
    assert bool((still == 1 - hit).all())
    sample = rng.integers(0, B, 300)
    np.testing.assert_array_equal(
        long_ok.cpu().numpy()[sample],
        Oracle.validate_control(data, res, -100.0, -100.0, (0.2, 1.0, 0.05, 0.65), x0.cpu().numpy()[sample],
                                u.cpu().numpy()[sample], 0.1, 2.0))


def test_integrate_twist_matches_cpu():
    """the constant-twist step + wrap (numerics.hpp:273-298, 77-89), both branches; sin / cos differ from
    glibc's by an ulp or two, everything else is rounded as the reference rounds it"""
    import torch

    import ergodic_exploration_b200 as eb

    rng = np.random.default_rng(8)
    n = 5000
    x = np.column_stack([rng.uniform(-50, 50, n), rng.uniform(-50, 50, n), rng.uniform(-np.pi, np.pi, n)])
    u = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-2, 2, n)])
    u[::4, 2] = 0.0
    got = eb.integrate_twist(torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda(), 0.1).cpu().numpy()
    want = np.empty_like(x)
    for i in range(n):
        w = Oracle.integrate_twist(x[i], u[i], 0.1)
        w[2] = Oracle.normalize_angle_pi(w[2])
        want[i] = w
    assert np.max(np.abs(got[:, :2] - want[:, :2])) <= 1e-13
    d = got[:, 2] - want[:, 2]
    assert np.max(np.abs(np.arctan2(np.sin(d), np.cos(d)))) <= 1e-13
    xd = torch.from_numpy(x).cuda()
    eb.integrate_twist(xd, torch.from_numpy(u).cuda(), 0.1, out=xd)  # in place
    np.testing.assert_array_equal(xd.cpu().numpy(), got)
