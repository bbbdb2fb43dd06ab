"""The synthetic C3 density (BASELINE.json configs[2]): an un-normalised mixture of 8 Gaussians on
the configTarget grid, mu ~ U([0.1 L, 0.9 L]^2), sigma ~ U(0.02 L, 0.10 L) per axis, parameters from a fixed numpy
stream so that the GPU benchmark, the GPU test and the CPU golden generator (tests/golden/make_golden_c3.py)
evaluate the SAME closed form (float64 exp on either side)."""
import numpy as np

SEED = 0xE16C0D1C + 3


def c3_params(n=8192, res=0.1):
    L = (n - 1) * res
    rng = np.random.default_rng(SEED)
    mu = (0.1 + 0.8 * rng.random((8, 2))) * L
    sg = (0.02 + 0.08 * rng.random((8, 2))) * L
    return mu, sg


def c3_density_numpy(n=8192, res=0.1, lo=0, hi=None):
    """rows lo..hi of the (n, n) density, x fastest"""
    hi = n if hi is None else hi
    mu, sg = c3_params(n, res)
    xs = np.arange(n, dtype=np.float64) * res
    ys = np.arange(lo, hi, dtype=np.float64) * res
    phi = np.zeros((hi - lo, n))
    for (mx, my), (sx, sy) in zip(mu, sg):
        phi += np.exp(-0.5 * ((xs[None, :] - mx) / sx) ** 2 - 0.5 * ((ys[:, None] - my) / sy) ** 2)
    return phi


def c3_density_torch(device, n=8192, res=0.1, lo=0, hi=None):
    import torch

    hi = n if hi is None else hi
    mu, sg = c3_params(n, res)
    xs = torch.arange(n, device=device, dtype=torch.float64) * res
    ys = torch.arange(lo, hi, device=device, dtype=torch.float64) * res
    phi = torch.zeros((hi - lo, n), dtype=torch.float64, device=device)
    for (mx, my), (sx, sy) in zip(mu, sg):
        phi += torch.exp(-0.5 * ((xs[None, :] - float(mx)) / float(sx)) ** 2 - 0.5 * ((ys[:, None] - float(my)) / float(sy)) ** 2)
    return phi
