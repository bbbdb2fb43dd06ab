"""-m gpu: the CUDA collision / DynamicWindow path (through the C ABI) against the committed golden
vectors generated from the compiled reference (tests/golden/make_golden_avoid.py)."""
import numpy as np
import pytest

from test_golden_avoid import load

pytestmark = pytest.mark.gpu


def test_cuda_path_against_golden():
    import ergodic_exploration_b200 as eb

    g, (data, res, xmin, ymin, col) = load()
    ys, xs = data.shape
    grid, c = eb.GridMap(xmin, xmin + xs * res, ymin, ymin + ys * res, res, data), eb.Collision(*col)
    for mode in (1, 2):  # circle walks, pre-dilated map
        grid.dilation(mode)
        np.testing.assert_array_equal(c.collisionCheck(grid, g["poses"]), g["hit"])
        np.testing.assert_array_equal(eb.validate_control(c, grid, g["inside"], g["twists"], 0.1, 0.5), g["valid_05"])
        np.testing.assert_array_equal(eb.validate_control(c, grid, g["inside"], g["twists"], 0.1, 2.0), g["valid_20"])
    grid.dilation(0)
    dwa = eb.DynamicWindow(c, *[float(v) for v in g["dwa_cfg"]], *[int(v) for v in g["samples"]])
    f, u = dwa.control(grid, g["inside"], g["twists"], vref=g["vref"])
    np.testing.assert_array_equal(f, g["dwa_found_twist"])
    np.testing.assert_array_equal(u, g["dwa_u_twist"])
    f, u = dwa.control(grid, g["inside"], g["twists"], xt_ref=g["xt_ref"], dt_ref=0.1)
    np.testing.assert_array_equal(f, g["dwa_found_traj"])
    np.testing.assert_array_equal(u, g["dwa_u_traj"])
