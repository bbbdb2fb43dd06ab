#!/usr/bin/env python
"""Where the end-to-end time of one C2 control() call goes (host clock, idle GPU before every call):
launch-path time of the device entry point, time to completion, and the host entry point as a whole."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import ergodic_exploration_b200 as eb  # noqa: E402

wl = bench.WORKLOADS["c2"]
B = wl["batch"]
R, umin, umax = bench.model_params(wl["model"])
x, ut, mem = bench.synth_inputs(wl, B, seed=3)
ctl = eb.ErgodicControl(wl["model"], bench.DT, wl["horizon"], 0.1, 1.0, wl["nb"], 1000, 100, R, umin, umax, batch=B)
ctl.setTarget([eb.Gaussian(m, s) for m, s in zip(bench.MU, bench.SIGMA)])
ctl.set_ut(ut)
ctl.keep_ck(False)
xd = torch.from_numpy(x).cuda()
u0d = torch.empty((B, 3), dtype=torch.float64, device="cuda")
xh = torch.from_numpy(x).pin_memory().numpy()
u0h = torch.empty((B, 3), dtype=torch.float64).pin_memory().numpy()
pc = time.perf_counter
for _ in range(5):
    ctl.control(bench.BOUNDS, xd, u0=u0d)
    ctl.control(bench.BOUNDS, xh, u0=u0h)
torch.cuda.synchronize()
n = 200
la, co, ho, em = [], [], [], []
s = torch.cuda.current_stream()
for _ in range(n):
    t0 = pc(); ctl.control(bench.BOUNDS, xd, u0=u0d); t1 = pc(); s.synchronize(); t2 = pc()
    la.append(t1 - t0); co.append(t2 - t1)
    t0 = pc(); ctl.control(bench.BOUNDS, xh, u0=u0h); t1 = pc()
    ho.append(t1 - t0)
    t0 = pc(); s.synchronize(); t1 = pc()
    em.append(t1 - t0)
med = lambda v: 1e6 * float(np.median(v))
print(f"device entry point: launch path {med(la):.1f} us, then until synchronize returns {med(co):.1f} us  (sum {med(la) + med(co):.1f})")
print(f"host entry point (zero-copy x / u0, fault flag, sync inside): {med(ho):.1f} us")
print(f"synchronize of an idle stream: {med(em):.2f} us")

# the C entry points alone (ctypes call with prebuilt arguments), idle GPU before every call
import ctypes as C  # noqa: E402
from ergodic_exploration_b200 import capi  # noqa: E402

lib = capi.load()
h = ctl._h
b = [C.c_double(v) for v in bench.BOUNDS]
xp, up = C.c_void_p(xd.data_ptr()), C.c_void_p(u0d.data_ptr())
xhp, uhp = C.c_void_p(xh.ctypes.data), C.c_void_p(u0h.ctypes.data)
cd, ch, st = [], [], []
for _ in range(n):
    t0 = pc(); lib.eb_control_dev(h, *b, xp, None, up, None); t1 = pc(); s.synchronize()
    cd.append(t1 - t0)
    t0 = pc(); lib.eb_control_host(h, *b, xhp, None, uhp, None); t1 = pc()
    ch.append(t1 - t0)
    t0 = pc(); lib.eb_steps(h); t1 = pc()
    st.append(t1 - t0)
print(f"ctypes -> eb_control_dev alone {med(cd):.1f} us; ctypes -> eb_control_host alone {med(ch):.1f} us; trivial ctypes call {med(st):.2f} us")

# zero-copy cost split: eb_control_dev with x and / or u0 in page-locked host memory (UVA: the host pointer is valid on the device)
for name, xa, ua in (("x dev,  u0 dev ", xp, up), ("x host, u0 dev ", xhp, up), ("x dev,  u0 host", xp, uhp), ("x host, u0 host", xhp, uhp)):
    tt = []
    for _ in range(n):
        t0 = pc(); lib.eb_control_dev(h, *b, xa, None, ua, None); s.synchronize(); t1 = pc()
        tt.append(t1 - t0)
    print(f"eb_control_dev + synchronize, {name}: {med(tt):.1f} us")
