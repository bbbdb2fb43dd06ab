"""-m gpu: map-derived target: int8 occupancy grid -> entropy density (numerics.hpp:164-179
over GridMap::getCell) -> phi_k, against the golden entropy values of the compiled reference and the CPU oracle's
spatialCoeff arithmetic at the cell centres."""
import os

import numpy as np
import pytest

from helpers import assert_abs_rel_close, assert_coeff_close
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_entropy.npz"))


def centre_phik(density, res, nb):
    """Basis::spatialCoeff on the normalised density at the cell centres (basis.cpp:122-133, target.cpp:87)"""
    ny, nx = density.shape
    lx, ly = nx * res, ny * res
    xs = 0.5 * res + np.cumsum(np.r_[0.0, np.full(nx - 1, res)])
    ys = 0.5 * res + np.cumsum(np.r_[0.0, np.full(ny - 1, res)])
    grid = np.stack([np.tile(xs, ny), np.repeat(ys, nx)], axis=1)
    vals = (density / density.sum()).reshape(-1)
    return Oracle.spatial_coeff(lx, ly, nb, vals, grid)


@pytest.mark.parametrize("nx,ny", [(96, 80), (640, 512)])
def test_map_target_wide_basis(nx, ny):
    """num_basis > 32: entropy density kernel + the wide phi_k routes (simple pair on the small grid, the TMA tile kernel
    per block of 32 x 32 orders on the large one)"""
    import ergodic_exploration_b200 as eb

    nb = 40
    rng = np.random.default_rng(nx)
    cells = rng.integers(-1, 101, size=(ny, nx)).astype(np.int8)
    mt = eb.MapTarget(nx, ny, 0.05, nb)
    got = mt.execute(cells)
    dens = Oracle.entropy_grid(cells)
    # eo_phik_rows takes accumulated sample points from 0: compare through the plain plan at the cell centres instead
    plan = eb.PhikPlan(nx, ny, 0.05, mt.lx, mt.ly, nb, x_first=0.025, y_first=0.025, algo=1)
    import torch
    simple = plan.execute(torch.from_numpy(dens).cuda()).cpu().numpy()
    assert_coeff_close(got, simple, f"map target nb=40 {nx}x{ny} vs the simple pair on the oracle's density")
    if nx * ny <= 10000:
        assert_coeff_close(got, centre_phik(dens, 0.05, nb), "map target nb=40 vs the oracle")


def test_entropy_density_matches_reference_golden():
    import torch

    import ergodic_exploration_b200 as eb

    cells = G["cells"]
    ny, nx = cells.shape
    mt = eb.MapTarget(nx, ny, 0.05, 8)
    phik = mt.execute(torch.from_numpy(cells).cuda())
    dens = mt.density().cpu().numpy()
    assert_abs_rel_close(dens, G["entropy"], "entropy density")  # CUDA log vs glibc log: last ulp
    assert dens[0, 0] == 1e-3 and dens[0, 1] == 1e-3 and dens[0, 2] == 0.7  # the three special cases are exact
    want = centre_phik(G["entropy"], 0.05, 8)
    assert_coeff_close(phik.cpu().numpy(), want, "phi_k of the entropy density")
    # host path, same numbers
    assert_coeff_close(mt.execute(cells), want, "host path")
    assert abs(mt.last_sum - G["entropy"].sum()) <= 1e-9 * G["entropy"].sum()


@pytest.mark.parametrize("nx,ny,nb", [(512, 512, 16), (1024, 1000, 20), (2048, 130, 32), (4096, 64, 8), (1000, 600, 10),
                                      (333, 257, 12)])
def test_map_target_shapes(nx, ny, nb):
    """large grids whose width is a multiple of 16 go through the FUSED kernel (TMA-staged bytes, entropy table in shared
    memory, folded DMMA tiles: the density is never written), other large grids through entropy kernel + folded TMA tile
    kernel, odd widths through the simple pair; compared with the oracle's entropy + spatialCoeff"""
    import torch

    import ergodic_exploration_b200 as eb

    rng = np.random.default_rng(nx + ny)
    cells = rng.integers(-1, 101, size=(ny, nx)).astype(np.int8)
    mt = eb.MapTarget(nx, ny, 0.05, nb)
    got = mt.execute(cells)
    dens = Oracle.entropy_grid(cells)
    want = centre_phik(dens, 0.05, nb)
    assert_coeff_close(got, want, f"map target {nx}x{ny} nb={nb}")
    # the density is materialised on demand from the last execute's cells
    assert_abs_rel_close(mt.density().cpu().numpy(), dens, "density on demand")
    # a second map update through the device path: new cells, same object; one launch when fused
    cells2 = np.where(rng.random((ny, nx)) < 0.7, 0, cells).astype(np.int8)
    l0 = mt.launch_count()
    got2 = mt.execute(torch.from_numpy(cells2).cuda()).cpu().numpy()
    fused = nx % 16 == 0 and nx * ny >= (1 << 18)
    assert mt.launch_count() - l0 == (1 if fused else (2 if nx * ny >= (1 << 18) and nx % 2 == 0 else 4))
    assert_coeff_close(got2, centre_phik(Oracle.entropy_grid(cells2), 0.05, nb), "second update")
    # the plain two-step route on the same density agrees to rounding
    plan = eb.PhikPlan(nx, ny, 0.05, mt.lx, mt.ly, nb, x_first=0.025, y_first=0.025)
    two_step = plan.execute(torch.from_numpy(Oracle.entropy_grid(cells2)).cuda()).cpu().numpy()
    assert_coeff_close(got2, two_step, "fused vs density + phi_k")


def test_controller_takes_the_map_target_without_host_round_trip():
    import torch

    import ergodic_exploration_b200 as eb
    from helpers import make_gpu, make_oracle, model_params, random_states, warm_ut
    from oracle.pyoracle import MODEL_OMNI

    rng = np.random.default_rng(11)
    nx = ny = 200
    res, nb, B = 0.05, 10, 32
    cells = rng.integers(-1, 101, size=(ny, nx)).astype(np.int8)
    mt = eb.MapTarget(nx, ny, res, nb)
    phik_dev = mt.execute(torch.from_numpy(cells).cuda())
    gpu = make_gpu(MODEL_OMNI, B, nb=nb)
    gpu.set_phik(phik_dev, mt.lx, mt.ly)  # device pointer in, no copy through the host
    x = random_states(rng, B)
    ut = warm_ut(rng, B, gpu.steps, MODEL_OMNI)
    gpu.set_ut(ut)
    u0 = gpu.control((0.0, mt.lx, 0.0, mt.ly), x)
    want_phik = centre_phik(Oracle.entropy_grid(cells), res, nb)
    for i in range(0, B, 5):
        o = make_oracle(MODEL_OMNI, nb=nb)
        o.set_phik(want_phik, mt.lx, mt.ly)
        o.set_ut(ut[i])
        assert_abs_rel_close(u0[i], o.control((0.0, mt.lx, 0.0, mt.ly), x[i]), "u0 with a map-derived target")
