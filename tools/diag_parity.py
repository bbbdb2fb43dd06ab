#!/usr/bin/env python
"""Per-quantity parity errors of one control() against the CPU oracle (debug aid)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import BOUNDS_10, make_gpu, make_oracle, random_states, warm_ut  # noqa: E402

for model, nb, horizon, nmem in ((1, 10, 5.0, 0), (0, 10, 5.0, 0), (1, 16, 5.0, 40), (0, 20, 10.0, 0)):
    rng = np.random.default_rng(7)
    B = 6
    steps = int(abs(horizon / 0.1))
    gpu = make_gpu(model, B, nb=nb, horizon=horizon)
    orcs = [make_oracle(model, nb=nb, horizon=horizon) for _ in range(B)]
    ut = warm_ut(rng, B, steps, model)
    gpu.set_ut(ut)
    for i, o in enumerate(orcs):
        o.set_ut(ut[i])
    for _ in range(nmem):
        past = random_states(rng, B)
        gpu.addStateMemory(past)
        for i, o in enumerate(orcs):
            o.add_state_memory(past[i])
    x = random_states(rng, B)
    met = np.empty(B)
    u0 = gpu.control(BOUNDS_10, x, metric=met)
    ck, utg = gpu.get_ck(), gpu.get_ut()
    e = {"u0": 0.0, "ck": 0.0, "metric": 0.0, "ut": 0.0}
    for i, o in enumerate(orcs):
        ou = o.control(BOUNDS_10, x[i])
        last = o.last()
        e["u0"] = max(e["u0"], np.abs(ou - u0[i]).max())
        e["ck"] = max(e["ck"], np.abs(last["ck"] - ck[i]).max() / np.abs(last["ck"]).max())
        e["metric"] = max(e["metric"], abs(last["metric"] - met[i]) / max(1.0, abs(last["metric"])))
        d = np.abs(o.get_ut() - utg[i])
        e["ut"] = max(e["ut"], d.max())
        if i == 0 and d.max() > 1e-9:
            print("   ut err per step (inst 0):", np.array2string(d.max(axis=1), precision=1, max_line_width=200))
            bad = np.abs(last["ck"] - ck[i]).reshape(nb, nb)
            if bad.max() > 1e-9:
                print("   ck err [ky][kx]:\n", np.array2string(bad, precision=1, max_line_width=250))
    print(f"model {model} nb {nb} N {steps} mem {nmem}: " + "  ".join(f"{k} {v:.2e}" for k, v in e.items()))
