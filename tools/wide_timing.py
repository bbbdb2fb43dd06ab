"""times control() for num_basis > 32 (solve_kernel_big) and the widest tuned kernel beside it"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ergodic_exploration_b200 as eb

R, umin, umax = np.diag([1.0, 1.0, 2.0]), [-1.0, -1.0, -2.0], [1.0, 1.0, 2.0]
peak = None
for nb, B, M in ((32, 16384, 0), (33, 16384, 0), (48, 8192, 0), (64, 4096, 0), (64, 4096, 100), (128, 1024, 0)):
    rng = np.random.default_rng(nb)
    ctl = eb.ErgodicControl(eb.Omni(), 0.1, 5.0, 0.1, 1.0, nb, 1000, 100, R, umin, umax, batch=B)
    ctl.setTarget([eb.Gaussian([2.5, 2.5], [1.5, 1.5]), eb.Gaussian([8.5, 2.5], [1.5, 1.5])])
    ctl.keep_ck(False)
    x = np.column_stack([rng.uniform(0.5, 9.5, B), rng.uniform(0.5, 9.5, B), rng.uniform(-np.pi, np.pi, B)])
    ctl.set_ut(rng.uniform(umin, umax, size=(B, ctl.steps, 3)) * 0.5)
    for _ in range(M):
        ctl.addStateMemory(np.column_stack([rng.uniform(0.5, 9.5, B), rng.uniform(0.5, 9.5, B), rng.uniform(-3, 3, B)]))
    xd = torch.from_numpy(x).cuda()
    u0 = torch.empty((B, 3), dtype=torch.float64, device="cuda")
    for _ in range(3):
        ctl.control((0.0, 10.0, 0.0, 10.0), xd, u0=u0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    a.record()
    for _ in range(K):
        ctl.control((0.0, 10.0, 0.0, 10.0), xd, u0=u0)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    N, Kc = ctl.steps, nb * nb
    F = 2 * Kc * (N + M) + 4 * Kc * N + 2 * Kc + 200 * N
    print(f"nb {nb:4d} B {B:6d} M {M:4d}: {ms:8.3f} ms/step, {B / ms * 1e3:10.3e} solves/s, {F * B / ms * 1e-9:7.2f} TFLOP/s algorithmic")
