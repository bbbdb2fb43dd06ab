#!/usr/bin/env python
"""How often can the last ulp of sin / cos move a pose across a cell edge?  ("weak" item 3)

The collision kernels chain integrate_twist steps with explicitly rounded operations in the reference's order; the
only difference to the host is CUDA's sincos against glibc's.  This tool rolls N robots x 20 constant-twist steps
on the GPU (eb.integrate_twist, the kernel the closed loops use) and on the CPU (oracle, glibc), then counts
  * poses whose world2Grid cell (floor((p - pmin) / res), grid.cpp:143-160) differs between the two,
  * poses within 4 ulp of a cell edge on the CPU side (the only ones that COULD flip),
  * the largest ulp distance between the two pose chains.
    python tools/edge_study.py [robots=5000000]      (needs a GPU; ~1 min)"""
import ctypes as C
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ergodic_exploration_b200 as eb  # noqa: E402
from oracle import pyoracle  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

pyoracle.build()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
steps, dt, res, pmin = 20, 0.1, 0.05, -100.0
rng = np.random.default_rng(0xE16C0D1C + 50)
x0 = np.column_stack([rng.uniform(-98, 98, n), rng.uniform(-98, 98, n), rng.uniform(-np.pi, np.pi, n)])
u = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-2, 2, n)])
u[: n // 50, 2] = 0.0  # the straight-line branch of integrate_twist

cpu = np.empty((steps, n, 3))
lib = Oracle.lib()
lib.eo_integrate_twist_chain.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_longlong, C.c_int, C.c_void_p]
threads = os.cpu_count() or 1
per = (n + threads - 1) // threads
tmp = [None] * threads


def work(t):
    lo, hi = t * per, min(n, (t + 1) * per)
    if lo >= hi:
        return
    out = np.empty((steps, hi - lo, 3))
    lib.eo_integrate_twist_chain(np.ascontiguousarray(x0[lo:hi]).ctypes.data, np.ascontiguousarray(u[lo:hi]).ctypes.data, dt,
                                 hi - lo, steps, out.ctypes.data)
    cpu[:, lo:hi] = out


ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
[t.start() for t in ts]
[t.join() for t in ts]

dev = torch.device("cuda", 0)
xd, ud = torch.from_numpy(x0).to(dev), torch.from_numpy(u).to(dev)
flips = near = 0
max_ulp = 0.0
for k in range(steps):
    eb.integrate_twist(xd, ud, dt, out=xd)
    g = xd.cpu().numpy()
    c = cpu[k]
    for a in range(2):
        tc, tg = (c[:, a] - pmin) / res, (g[:, a] - pmin) / res
        flips += int(np.count_nonzero(np.floor(tc) != np.floor(tg)))
        near += int(np.count_nonzero(np.abs(tc - np.rint(tc)) <= 4 * np.spacing(np.abs(tc))))
        max_ulp = max(max_ulp, float(np.max(np.abs(c[:, a] - g[:, a])) / np.spacing(100.0)))
poses = steps * n
print(f"poses checked                      {poses} ({n} robots x {steps} steps, map origin {pmin}, resolution {res})")
print(f"cell index differs GPU vs CPU       {flips} coordinates")
print(f"CPU coordinates within 4 ulp of an edge  {near}")
print(f"largest GPU-CPU distance           {max_ulp:.1f} ulp of the map scale (ulp(100 m) = {np.spacing(100.0):.2e} m)")
t = 4000.0
print(f"expected edge-proximity rate       ~{2 * 8 * np.spacing(t) :.1e} per coordinate (8 ulp window at cell index ~{t:.0f}) "
      f"-> {2 * poses * 8 * np.spacing(t):.1e} expected in this sample")
