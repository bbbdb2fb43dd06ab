#!/usr/bin/env python
"""Golden vectors for the widened rows of round 2, generated through the COMPILED REFERENCE (oracle/_ref: the unmodified
reference sources over the test-only Armadillo shim): forward RK4 rollouts of models::Cart and models::Mecanum
(integrator.hpp:135-152), f / fdx / fdu / wheels2Twist of both, and entropy() of a GridMap's cells
(numerics.hpp:164-179, grid.cpp:177-184).  Only possible where /root/reference exists.

    python tests/golden/make_golden_models.py  ->  tests/golden/models_entropy.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402


def main():
    pyoracle.build()
    assert RefLib.available(), "compiled reference not available"
    rng = np.random.default_rng(0xE16C0D1C + 21)
    n, steps, dt = 24, 30, 0.1
    out = {"dt": dt, "steps": steps}
    x0 = np.column_stack([rng.uniform(-3, 3, n), rng.uniform(-3, 3, n), rng.uniform(-np.pi, np.pi, n)])
    out["x0"] = x0
    cart_p, mec_p = (0.1, 2.0), (0.05, 0.3, 0.2)
    out["cart_params"], out["mecanum_params"] = np.array(cart_p), np.array(mec_p)
    ut_c = rng.uniform(-20, 20, (n, steps, 2))
    ut_m = rng.uniform(-30, 30, (n, steps, 4))
    out["ut_cart"], out["ut_mecanum"] = ut_c, ut_m
    out["xt_cart"] = np.stack([RefLib.rk4_forward_cart(*cart_p, dt, steps * dt, x0[i], ut_c[i]) for i in range(n)])
    out["xt_mecanum"] = np.stack([RefLib.rk4_forward_mecanum(*mec_p, dt, steps * dt, x0[i], ut_m[i]) for i in range(n)])
    fc = [RefLib.cart(*cart_p, x0[i], ut_c[i, 0]) for i in range(n)]
    fm = [RefLib.mecanum(*mec_p, x0[i], ut_m[i, 0]) for i in range(n)]
    for name, res in (("cart", fc), ("mecanum", fm)):
        for j, key in enumerate(("f", "A", "B", "vb")):
            out[f"{name}_{key}"] = np.stack([np.asarray(r[j]) for r in res])
    # the reference's own integrator test (test/test_integrator.cpp:44-73): Cart(0.1, 2.0), u = [1, 1], 4 steps
    out["kat_xt"] = RefLib.rk4_forward_cart(0.1, 2.0, 0.1, 0.4, [0.0, 0.0, 0.0], np.ones((4, 2)))
    cells = rng.integers(-1, 101, size=(96, 160)).astype(np.int8)
    cells[0, :8] = [0, 100, -1, 50, 1, 99, 37, -1]
    out["cells"] = cells
    out["entropy"] = RefLib.entropy_grid(cells)
    path = os.path.join(ROOT, "tests", "golden", "models_entropy.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
