#!/usr/bin/env python
"""Kernel-only timing of the fused solve kernel for tuning variants (no CPU baseline, no e2e):

    EB_LIB_PATH=variants/lib_x.so python tools/ktime.py [c2 c5 c4 ...]

prints one line per workload: ms per launch (CUDA events, L2 flushed for small working sets) and the FP64 fraction."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import ergodic_exploration_b200 as eb  # noqa: E402

names = sys.argv[1:] or ["c2", "c5", "c4"]
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tag = os.path.basename(os.environ.get("EB_LIB_PATH", "libergodic_b200.so"))
for key in names:
    wl = bench.WORKLOADS[key]
    B = int(os.environ.get("EB_KTIME_BATCH", wl["batch"]))
    R, umin, umax = bench.model_params(wl["model"])
    N, K, M = int(abs(wl["horizon"] / bench.DT)), wl["nb"] ** 2, min(wl["mem"], 100)
    x, ut, mem = bench.synth_inputs(wl, B, seed=0xE16C0D1C + 2)
    ctl = eb.ErgodicControl(wl["model"], bench.DT, wl["horizon"], 0.1, 1.0, wl["nb"], 1000000, 100, R, umin, umax, batch=B)
    ctl.setTarget([eb.Gaussian(m, s) for m, s in zip(bench.MU, bench.SIGMA)])
    ctl.set_ut(ut)
    ctl.keep_ck(False)
    if mem is not None:
        for m in mem:
            ctl.addStateMemory(m)
    xd = torch.from_numpy(x).to(dev)
    u0 = torch.empty((B, 3), dtype=torch.float64, device=dev)
    met = torch.empty(B, dtype=torch.float64, device=dev)
    for _ in range(5):
        ctl.control(bench.BOUNDS, xd, u0=u0, metric=met)
    torch.cuda.synchronize()
    steps = 20
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda._sleep(int(steps * 150e-6 * 1.9e9))
    for a, b in ev:
        flush.zero_()
        a.record()
        ctl.control(bench.BOUNDS, xd, u0=u0, metric=met)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    ms = sum(ts) / steps
    F = bench.flops_per_solve(K, N, M)
    print(f"{tag:28s} {key:6s} B={B:7d}  mean {ms:8.4f} ms  min {ts[0]:8.4f}  frac(37.2) {F * B / (ms * 1e-3) / 37.2e12:5.3f}  "
          f"u0sum {float(u0.sum()):+.6e}", flush=True)
    ctl.close()
