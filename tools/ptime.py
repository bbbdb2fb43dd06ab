#!/usr/bin/env python
"""phi_k kernels side by side: timing (CUDA events) and agreement of every algorithm on one density.

    python tools/ptime.py [n=8192] [nb=32]
algo 1 simple, 2 / 3 register-streamed DMMA tiles (fold / no fold), 4 / 5 TMA-staged DMMA tiles (fold / no fold)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ergodic_exploration_b200 as eb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ny = int(sys.argv[3]) if len(sys.argv) > 3 else n
res = 0.1
lx, ly = (n - 1) * res, (ny - 1) * res
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
phi = torch.rand((ny, n), generator=g, device=dev, dtype=torch.float64)
ref = None
ALGOS = [int(a) for a in os.environ.get("EB_PTIME_ALGOS", "1,2,3,4,5").split(",")]
for algo in ALGOS:
    try:
        plan = eb.PhikPlan(n, ny, res, lx, ly, nb, algo=algo)
    except Exception as exc:
        print(f"algo {algo}: unsupported ({exc})")
        continue
    out = torch.empty(nb * nb, dtype=torch.float64, device=dev)
    for _ in range(3):
        plan.execute(phi, out)
    torch.cuda.synchronize()
    steps = 5 if algo == 1 else 30
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record()
        plan.execute(phi, out)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        plan.execute(phi, out)
    b.record()
    torch.cuda.synchronize()
    one_pair = a.elapsed_time(b) / steps
    got = out.cpu().numpy()
    if ref is None:
        ref = got
    err = abs(got - ref).max() / abs(ref).max()
    print(f"algo {algo}: {n}x{ny} nb {nb}  mean {sum(ms) / steps:8.4f} ms  min {ms[0]:8.4f}  HBM {8.0 * n * ny / (sum(ms) / steps * 1e-3) / 1e9:7.1f} GB/s"
          f"  one pair around {steps}: {one_pair:8.4f} ms/step  fold {plan.fold()[0]}  max rel diff vs first {err:.2e}", flush=True)
    plan.close()
