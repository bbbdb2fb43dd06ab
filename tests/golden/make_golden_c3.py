#!/usr/bin/env python
"""Golden phi_k of the FULL C3 config (BASELINE.json configs[2]: 8192 x 8192 density, 32 x 32 basis) from the CPU oracle.

    python tests/golden/make_golden_c3.py        (about a minute on 8 cores; writes tests/golden/c3_phik_8192.npz)

The density is tools/c3_density.py (closed form, fixed numpy stream).  Every cell goes through the arithmetic of
Basis::spatialCoeff (basis.cpp:122-133) on Target::fill's normalised density (target.cpp:87) as restated by
oracle/ergodic_oracle.c::eo_phik_rows -- accumulated grid coordinates, cos(k (PI / l) x) per axis, F_k * (phi / total)
summed in row-major order -- with the rows split into blocks that run on separate threads; the blocks' partial
sums are then added in block order.  (The literal reference cannot run this size: its K x G temporary is 550 TB.)
Also stored: a 1024 x 768 case with nb = 20 (ragged, no mirror fold) from the single-threaded eo_phik_from_grid."""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from c3_density import c3_density_numpy  # noqa: E402

from oracle import pyoracle  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402


def main():
    pyoracle.build()
    n, nb, res = 8192, 32, 0.1
    L = (n - 1) * res
    t0 = time.time()
    phi = c3_density_numpy(n, res)
    total = float(phi.sum())  # numpy pairwise sum; the GPU normalises by its own raw[0][0] (differs by ~1e-16 relative)
    blocks = 64
    rows = n // blocks
    parts = [None] * blocks
    threads = os.cpu_count() or 1

    def work(tid):
        for b in range(tid, blocks, threads):
            parts[b] = Oracle.phik_rows(phi[b * rows:(b + 1) * rows], b * rows, res, L, L, nb, total)

    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    phik = np.zeros(nb * nb)
    for b in range(blocks):
        phik += parts[b]
    print(f"8192^2, nb 32: {time.time() - t0:.1f} s, phik[0] = {phik[0]:.17g}, sum(phi) = {total:.17g}")

    # a ragged, asymmetric case through the literal single-threaded restatement
    rng = np.random.default_rng(0xE16C0D1C + 33)
    small = rng.random((768, 1024))
    lx2, ly2 = 1023 * res, 767 * res
    ph2, s2 = Oracle.phik_from_grid(small, res, lx2, ly2, 20)
    out = os.path.join(ROOT, "tests", "golden", "c3_phik_8192.npz")
    np.savez_compressed(out, phik=phik, phi_sum=total, n=n, nb=nb, res=res, blocks=blocks,
                        small_seed=0xE16C0D1C + 33, small_phik=ph2, small_sum=s2)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
