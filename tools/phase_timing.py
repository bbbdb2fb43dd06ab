#!/usr/bin/env python
"""Where the wall time of one solve_kernel launch goes: builds a debug copy of
the library with -DEB_PHASE_TIMING (clock64() stamps at the phase boundaries of
every instance), runs one bench workload and prints, per phase, the mean
per-warp duration and the span over the whole grid.

    python tools/phase_timing.py [c2|c5|c4|c2big]
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DBG = os.path.join(ROOT, "variants", "lib_phase.so")  # tools/variants.sh phase "-DEB_PHASE_TIMING"
if not os.path.exists(DBG):
    csrc = os.path.join(ROOT, "ergodic_exploration_b200", "csrc")
    os.makedirs(os.path.dirname(DBG), exist_ok=True)
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-Xcompiler", "-fPIC", "-DEB_PHASE_TIMING", "-shared", "-o", DBG,
                           os.path.join(csrc, "ergodic_b200.cu"), "-lcudart"])
os.environ["EB_LIB_PATH"] = DBG

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import ergodic_exploration_b200 as eb  # noqa: E402
from ergodic_exploration_b200 import capi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WORKLOADS[name]
B = min(wl["batch"], 65536)
R, umin, umax = bench.model_params(wl["model"])
x, ut, mem = bench.synth_inputs(wl, B, seed=1)
ctl = eb.ErgodicControl(wl["model"], bench.DT, wl["horizon"], 0.1, 1.0, wl["nb"], 1000000, 100, R, umin, umax, batch=B)
ctl.setTarget([eb.Gaussian(m, s) for m, s in zip(bench.MU, bench.SIGMA)])
ctl.set_ut(ut)
if mem is not None:
    for m in mem:
        ctl.addStateMemory(m)
xd = torch.from_numpy(x).cuda()
for _ in range(3):
    ctl.control(bench.BOUNDS, xd)
torch.cuda.synchronize()
lib = capi.load()
S = 16
buf = np.zeros((B, S), dtype=np.int64)
assert lib.eb_debug_phase_dump(buf.ctypes.data_as(C.c_void_p), B) == 0
names = ["replay c_k", "round 0: rollout + trig", "round 0: c_k tables + DMMA", "round 1: rollout + trig",
         "round 1: c_k tables + DMMA", "S / metric", "last round: gradient", "last round: co-state + update",
         "first round: gradient", "first round: co-state + update", "exit"]
sm = buf[:, 15]
print(f"workload {name}: {B} instances (solve_kernel v1 stamps; horizons of at most 2 rounds); cycles of the SM clock")
spans = []
for s in np.unique(sm):
    t = buf[sm == s]
    spans.append((t[:, 11].max() - t[:, 0].min(), len(t)))
spans = np.array(spans)
print(f"SM busy span: mean {spans[:, 0].mean():.0f} cycles, max {spans[:, 0].max():.0f}; warps per SM {spans[:, 1].mean():.1f}")
tot = (buf[:, 11] - buf[:, 0]).mean()
for i, n in enumerate(names):
    d = buf[:, i + 1] - buf[:, i]
    print(f"  {n:32s} mean {d.mean():9.0f} cycles  ({100 * d.mean() / tot:5.1f}% of a warp's {tot:.0f})  min {d.min()} max {d.max()}")
