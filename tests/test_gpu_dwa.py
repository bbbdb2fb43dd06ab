"""-m gpu: batched DynamicWindow::control on the GPU (through the C ABI) against the CPU
restatement: the selected twist is index work (identical, up to the rounding of the
accumulated candidate twists which is reproduced), found flags identical, costs to 1e-9."""
import numpy as np
import pytest

from helpers import assert_abs_rel_close
from oracle.pyoracle import Oracle
from test_dwa_cpu import COL, DWA_CART, DWA_OMNI, scenario

pytestmark = pytest.mark.gpu


def _setup(data, res):
    import ergodic_exploration_b200 as eb

    ys, xs = data.shape
    return eb.GridMap(0.0, xs * res, 0.0, ys * res, res, data), eb.Collision(*COL)


@pytest.mark.parametrize("cfg,samples", [(DWA_OMNI, (3, 8, 5)), (DWA_CART, (5, 1, 9)), (DWA_OMNI, (0, 2, 1)),
                                         (DWA_OMNI, (7, 7, 7))])
def test_dwa_twist_matches_cpu(cfg, samples):
    import ergodic_exploration_b200 as eb

    rng = np.random.default_rng(3 + sum(samples))
    n = 600
    data, res, x0, vb = scenario(rng, n)
    vref = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-2, 2, n)])
    grid, col = _setup(data, res)
    dwa = eb.DynamicWindow(col, *cfg, *samples)
    cost = np.empty(n)
    fo, uo, co = Oracle.dwa_control(data, res, 0.0, 0.0, COL, cfg, samples, x0, vb, vref=vref)
    for mode in (1, 2):  # circle walks, pre-dilated map
        grid.dilation(mode)
        found, u = dwa.control(grid, x0, vb, vref=vref, min_cost=cost)
        np.testing.assert_array_equal(found, fo)
        np.testing.assert_array_equal(u, uo)
    assert_abs_rel_close(cost[fo == 1], co[fo == 1], "min cost")
    assert 0 < fo.sum()


@pytest.mark.parametrize("per_instance", [False, True])
def test_dwa_trajectory_matches_cpu(per_instance):
    import ergodic_exploration_b200 as eb

    rng = np.random.default_rng(29)
    n, cfg, samples = 300, DWA_OMNI, (3, 8, 5)
    data, res, x0, vb = scenario(rng, n)
    t = np.arange(50) * 0.1
    base = np.column_stack([4.0 + 1.5 * np.cos(0.7 * t), 3.5 + 1.2 * np.sin(0.9 * t), 4.0 * np.sin(0.5 * t)])
    grid, col = _setup(data, res)
    dwa = eb.DynamicWindow(col, *cfg, *samples)
    if per_instance:
        xt = base[None] + rng.normal(0, 0.3, (n, 1, 3))
        found, u = dwa.control(grid, x0, vb, xt_ref=xt, dt_ref=0.1)
        for i in range(0, n, 7):
            fo, uo, _ = Oracle.dwa_control(data, res, 0.0, 0.0, COL, cfg, samples, x0[i:i + 1], vb[i:i + 1], xt_ref=xt[i],
                                           dt_ref=0.1)
            assert found[i] == fo[0]
            np.testing.assert_array_equal(u[i], uo[0])
    else:
        cost = np.empty(n)
        found, u = dwa.control(grid, x0, vb, xt_ref=base, dt_ref=0.1, min_cost=cost)
        fo, uo, co = Oracle.dwa_control(data, res, 0.0, 0.0, COL, cfg, samples, x0, vb, xt_ref=base, dt_ref=0.1)
        np.testing.assert_array_equal(found, fo)
        np.testing.assert_array_equal(u, uo)
        assert_abs_rel_close(cost[fo == 1], co[fo == 1], "min cost")


def test_dwa_device_path_and_no_solution():
    import torch

    import ergodic_exploration_b200 as eb

    rng = np.random.default_rng(5)
    n = 4096
    data, res, x0, vb = scenario(rng, n)
    vref = np.zeros((n, 3))
    grid, col = _setup(data, res)
    dwa = eb.DynamicWindow(col, *DWA_OMNI, 3, 8, 5)
    fh, uh = dwa.control(grid, x0, vb, vref=vref)
    fd, ud = dwa.control(grid, torch.from_numpy(x0).cuda(), torch.from_numpy(vb).cuda(), vref=torch.from_numpy(vref).cuda())
    np.testing.assert_array_equal(fh, fd.cpu().numpy())
    np.testing.assert_array_equal(uh, ud.cpu().numpy())
    grid.update(np.full(data.shape, 100, dtype=np.int8))  # nothing is collision free any more
    f2, u2 = dwa.control(grid, x0, vb, vref=vref)
    assert not f2.any() and not u2.any()
    assert dwa.steps() == 20 and dwa.timeStep() == 0.1
