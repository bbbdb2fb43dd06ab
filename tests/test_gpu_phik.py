"""-m gpu: phi_k target coefficients (Target::fill + Basis::spatialCoeff) on the
GPU against the CPU oracle.  Coefficient tolerance: max|a-b|/max|b| <= 1e-9."""
import numpy as np
import pytest

from helpers import assert_coeff_close, oracle_phik_threaded
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


def gaussian_mixture(rng, nx, ny, res, ng=8):
    """un-normalised mixture, mu ~ U([0.1L, 0.9L]^2), sigma ~ U(0.02L, 0.10L)"""
    lx, ly = (nx - 1) * res, (ny - 1) * res
    xs, ys = np.arange(nx) * res, np.arange(ny) * res
    phi = np.zeros((ny, nx))
    for _ in range(ng):
        mx, my = rng.uniform(0.1 * lx, 0.9 * lx), rng.uniform(0.1 * ly, 0.9 * ly)
        sx, sy = rng.uniform(0.02 * lx, 0.10 * lx), rng.uniform(0.02 * ly, 0.10 * ly)
        phi += np.exp(-0.5 * ((xs[None, :] - mx) / sx) ** 2 - 0.5 * ((ys[:, None] - my) / sy) ** 2)
    return phi, lx, ly


@pytest.mark.parametrize("nx,ny,nb,algo", [
    (101, 101, 10, 0),   # the shipped default: 10 m map @ 0.1, 10x10 basis
    (200, 319, 10, 1),   # maze extent
    (64, 48, 32, 1),
    (37, 29, 5, 1),      # ragged
    (3, 2, 3, 1),        # tiny
    (256, 128, 32, 0),
    (512, 384, 16, 0),
    (256, 100, 32, 2),   # DMMA/TMA tile kernel, ragged row block
    (132, 77, 10, 2),    # ragged column chunk (132 = 128 + 4)
    (1024, 520, 32, 2),
    (512, 64, 7, 2),
    (1280, 1100, 24, 0),  # auto -> tile kernel
    (1024, 520, 32, 3),   # tile kernel with the mirror fold switched off
    (136, 70, 32, 2),     # fold, nx / 2 = 68: ragged folded chunk
    (136, 70, 32, 3),
    (2056, 130, 31, 2),   # fold, several units per row block, odd basis count
    (101, 101, 33, 0),    # num_basis > 32: the simple pair over blocks of 32 orders
    (160, 90, 64, 1),
    (77, 130, 100, 0),
    (640, 512, 48, 0),    # large grid: the TMA tile kernel once per block of 32 x 32 orders (mirror-folded)
    (640, 512, 48, 1),    # ... and the simple pair on the same grid
    (1026, 300, 64, 0),   # 2 x 2 full blocks, nx / 2 odd
    (528, 500, 100, 0),   # 4 x 4 blocks, last one 4 orders wide
])
def test_phik_matches_oracle(nx, ny, nb, algo):
    from ergodic_exploration_b200 import PhikPlan

    rng = np.random.default_rng(nx * 1000 + ny)
    res = 0.1
    phi, lx, ly = gaussian_mixture(rng, nx, ny, res)
    lx, ly = max(lx, res), max(ly, res)
    plan = PhikPlan(nx, ny, res, lx, ly, nb, algo=algo)
    got = plan.execute(phi)
    want, total = (oracle_phik_threaded if nx * ny * nb * nb > 2e8 else Oracle.phik_from_grid)(phi, res, lx, ly, nb)
    assert_coeff_close(got, want, f"phi_k {nx}x{ny} nb={nb}")
    assert abs(plan.last_sum - total) <= 1e-9 * abs(total)
    assert abs(got[0] - 1.0) < 1e-12  # phi_0 == 1 after normalisation


def test_phik_arbitrary_density_not_separable():
    """the kernel must not rely on Gaussian separability: random dense density"""
    from ergodic_exploration_b200 import PhikPlan

    rng = np.random.default_rng(99)
    nx, ny, nb, res = 160, 96, 12, 0.25
    phi = rng.random((ny, nx))
    lx, ly = (nx - 1) * res, (ny - 1) * res
    got = PhikPlan(nx, ny, res, lx, ly, nb).execute(phi)
    want, _ = Oracle.phik_from_grid(phi, res, lx, ly, nb)
    assert_coeff_close(got, want, "random density")


def test_phik_mirror_fold_gate_and_bound():
    """the fold is taken only on grids whose cosine table is mirror-symmetric
    (lx == (nx - 1) res); folded and unfolded tile kernels agree to the measured
    table asymmetry; a grid with another lx falls back and still matches the oracle"""
    from ergodic_exploration_b200 import PhikPlan

    rng = np.random.default_rng(77)
    nx, ny, nb, res = 1024, 200, 32, 0.1
    phi = rng.random((ny, nx)) ** 3  # no symmetry of its own
    lx, ly = (nx - 1) * res, (ny - 1) * res
    folded, plain = PhikPlan(nx, ny, res, lx, ly, nb, algo=2), PhikPlan(nx, ny, res, lx, ly, nb, algo=3)
    ok, dev = folded.fold()
    assert ok and dev <= 1e-10
    a, b = folded.execute(phi), plain.execute(phi)
    assert np.max(np.abs(a - b)) <= max(4.0 * dev, 1e-13)
    want, _ = Oracle.phik_from_grid(phi, res, lx, ly, nb)
    assert_coeff_close(a, want, "folded")
    assert_coeff_close(b, want, "unfolded")
    # the reference's configTarget grid can overshoot: lx = 102.33 -> nx = round(lx / res) + 1 = 1024, last point 102.3
    lx2 = 102.33
    skew = PhikPlan(nx, ny, res, lx2, ly, nb, algo=2)
    ok2, dev2 = skew.fold()
    assert not ok2 and dev2 > 1e-10
    want2, _ = Oracle.phik_from_grid(phi, res, lx2, ly, nb)
    assert_coeff_close(skew.execute(phi), want2, "asymmetric grid -> unfolded")


def test_phik_linearity_large():
    """size-independent property at a size the oracle cannot reach in seconds:
    the un-normalised contraction is linear, phi_k(a+b) * sum(a+b) ==
    phi_k(a) * sum(a) + phi_k(b) * sum(b)"""
    from ergodic_exploration_b200 import PhikPlan

    rng = np.random.default_rng(2024)
    nx = ny = 2048
    nb, res = 32, 0.1
    a, lx, ly = gaussian_mixture(rng, nx, ny, res)
    b = rng.random((ny, nx))
    plan = PhikPlan(nx, ny, res, lx, ly, nb)
    pa = plan.execute(a); sa = plan.last_sum
    pb = plan.execute(b); sb = plan.last_sum
    pab = plan.execute(a + b); sab = plan.last_sum
    assert abs(sab - (sa + sb)) <= 1e-10 * sab
    assert_coeff_close(pab * sab, pa * sa + pb * sb, "linearity")


def test_phik_wide_raw_block_and_tile_algos_refused():
    """num_basis > 32: the raw block has leading dimension nb rounded up to 32; the 32-order tile kernels say no"""
    import torch

    from ergodic_exploration_b200 import ErgodicB200Error, PhikPlan, capi

    rng = np.random.default_rng(5)
    nx, ny, nb, res = 128, 64, 40, 0.1
    phi = rng.random((ny, nx))
    lx, ly = (nx - 1) * res, (ny - 1) * res
    plan = PhikPlan(nx, ny, res, lx, ly, nb)
    raw = torch.full((64, 64), np.nan, dtype=torch.float64, device="cuda")
    plan.execute_raw(torch.from_numpy(phi).cuda(), raw=raw)
    raw = raw.cpu().numpy()
    want, total = Oracle.phik_from_grid(phi, res, lx, ly, nb)
    assert_coeff_close(raw[:nb, :nb].ravel() / raw[0, 0], want, "raw block nb=40")
    assert abs(raw[0, 0] - total) <= 1e-9 * total
    assert np.all(raw[nb:, :] == 0.0) and np.all(raw[:, nb:] == 0.0)
    with pytest.raises(ErgodicB200Error) as e:
        PhikPlan(nx, ny, res, lx, ly, nb, algo=4)
    assert e.value.status == capi.EB_ERR_UNSUPPORTED
    # the same through the per-block TMA route (large grid): raw blocks assembled into the ld x ld layout
    nx, ny = 1024, 512
    phi = rng.random((ny, nx))
    lx, ly = (nx - 1) * res, (ny - 1) * res
    plan = PhikPlan(nx, ny, res, lx, ly, nb)
    raw = torch.full((64, 64), np.nan, dtype=torch.float64, device="cuda")
    plan.execute_raw(torch.from_numpy(phi).cuda(), raw=raw)
    raw = raw.cpu().numpy()
    want, total = oracle_phik_threaded(phi, res, lx, ly, nb)
    assert_coeff_close(raw[:nb, :nb].ravel() / raw[0, 0], want, "raw blocks nb=40, TMA route")
    assert np.all(raw[nb:, :] == 0.0) and np.all(raw[:, nb:] == 0.0)
    assert plan.launch_count() >= 5  # 4 tile-kernel launches + the assembly


def test_plan_rejects_bad_arguments():
    from ergodic_exploration_b200 import PhikPlan

    with pytest.raises(ValueError):
        PhikPlan(16, 16, 0.1, 1.5, 1.5, 129)  # EB_MAX_NUM_BASIS = 128
    with pytest.raises(ValueError):
        PhikPlan(0, 16, 0.1, 1.5, 1.5, 4)
