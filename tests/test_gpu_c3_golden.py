"""-m gpu: the FULL C3 config (BASELINE.json configs[2]: 8192 x 8192 density, 32 x 32 basis) against the committed
golden coefficients of the CPU oracle (tests/golden/c3_phik_8192.npz, made by tests/golden/make_golden_c3.py from the
same closed-form density, tools/c3_density.py), for every large-grid kernel: TMA-staged and register-streamed
DMMA tiles, each with and without the mirror fold.  Tolerance: max|a-b| / max|b| <= 1e-9."""
import os
import sys

import numpy as np
import pytest

from helpers import assert_coeff_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
GOLDEN = os.path.join(ROOT, "tests", "golden", "c3_phik_8192.npz")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    import torch
    from c3_density import c3_density_torch

    g = np.load(GOLDEN)
    n, res = int(g["n"]), float(g["res"])
    phi = c3_density_torch(torch.device("cuda", 0), n, res)
    yield g, phi, n, res
    del phi
    torch.cuda.empty_cache()


@pytest.mark.parametrize("algo,name", [(0, "auto"), (4, "TMA tiles, mirror fold"), (5, "TMA tiles, no fold"),
                                       (2, "register-streamed tiles, mirror fold"), (3, "register-streamed tiles, no fold")])
def test_c3_full_size_against_golden(c3, algo, name):
    import torch

    from ergodic_exploration_b200 import PhikPlan

    g, phi, n, res = c3
    nb = int(g["nb"])
    L = (n - 1) * res
    plan = PhikPlan(n, n, res, L, L, nb, algo=algo)
    fold, dev = plan.fold()
    assert fold and dev <= 1e-10  # the configTarget grid is mirror-symmetric to 1.4e-11 at this size
    out = torch.empty(nb * nb, dtype=torch.float64, device=phi.device)
    tot = torch.empty(1, dtype=torch.float64, device=phi.device)
    plan.execute(phi, out, tot)
    got = out.cpu().numpy()
    assert_coeff_close(got, g["phik"], f"C3 8192^2 nb=32, {name}")
    assert abs(float(tot) - float(g["phi_sum"])) <= 1e-9 * float(g["phi_sum"])
    assert abs(got[0] - 1.0) <= 1e-12
    # a second execute on the same plan gives the same bits (deterministic partial order, counter reset)
    out2 = torch.empty_like(out)
    plan.execute(phi, out2)
    assert torch.equal(out, out2)


def test_ragged_asymmetric_grid_against_golden():
    """1024 x 768, nb = 20, random dense density from the literal single-threaded restatement"""
    from ergodic_exploration_b200 import PhikPlan

    g = np.load(GOLDEN)
    rng = np.random.default_rng(int(g["small_seed"]))
    phi = rng.random((768, 1024))
    res = 0.1
    for algo in (0, 4, 5, 2):
        got = PhikPlan(1024, 768, res, 1023 * res, 767 * res, 20, algo=algo).execute(phi)
        assert_coeff_close(got, g["small_phik"], f"1024x768 nb=20 algo {algo}")


@pytest.mark.parametrize("nx,ny,nb", [(64, 64, 8), (66, 129, 32), (1030, 770, 20), (4096, 100, 32), (130, 2000, 5)])
def test_tma_kernel_shapes_against_oracle(nx, ny, nb):
    """odd shapes through the TMA kernel (even nx >= 64): ragged last band, ragged last chunk, narrow grids"""
    from oracle.pyoracle import Oracle

    from ergodic_exploration_b200 import PhikPlan

    rng = np.random.default_rng(nx * 7 + ny)
    phi = rng.random((ny, nx))
    res = 0.1
    lx, ly = (nx - 1) * res, (ny - 1) * res
    want, total = Oracle.phik_from_grid(phi, res, lx, ly, nb)
    for algo in (4, 5):
        plan = PhikPlan(nx, ny, res, lx, ly, nb, algo=algo)
        got = plan.execute(phi)
        assert_coeff_close(got, want, f"TMA phi_k {nx}x{ny} nb={nb} algo={algo}")
        assert abs(plan.last_sum - total) <= 1e-9 * abs(total)
