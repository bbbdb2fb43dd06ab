#!/usr/bin/env python
"""Per-CTA time line of the phi_k TMA kernel (needs a library built with -DEB_PT_TRACE, see tools/variants.sh):
    EB_LIB_PATH=variants/lib_pttrace.so python tools/phik_trace.py [ny ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ergodic_exploration_b200 as eb  # noqa: E402

n, nb, res = 8192, 32, 0.1
dev = torch.device("cuda", 0)
for ny in [int(a) for a in sys.argv[1:]] or [512, 8192]:
    phi = torch.rand((ny, n), device=dev, dtype=torch.float64)
    plan = eb.PhikPlan(n, ny, res, (n - 1) * res, (ny - 1) * res, nb, algo=4)
    out = torch.empty(nb * nb, dtype=torch.float64, device=dev)
    for _ in range(4):
        plan.execute(phi, out)
    torch.cuda.synchronize()
    plan.close()
