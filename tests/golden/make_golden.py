#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference sources compiled
in this container (oracle/_ref/libergodic_ref.so, built by `make -C oracle ref`
from /root/reference against the test-only Armadillo/ROS shim).

/root/reference does not exist on the GPU box, so the vectors are committed;
re-run this script here to regenerate them:

    python tests/golden/make_golden.py

Each case stores the full input state of one control() call (teacher-forced:
x, ut_ before, stored memory) and the reference's outputs (u0, ut_ after, c_k,
phi_k), so any implementation can be checked one step at a time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from helpers import (BOUNDS_10, BOUNDS_MAZE, MODEL_OMNI, MODEL_SIMPLE_CART, make_oracle, plant,  # noqa: E402
                     random_states, warm_ut)
from oracle.pyoracle import RefLib  # noqa: E402


def closed_loop_case(model, bounds, steps=6):
    """config C1: x0 = origin + (5, 7, 0.3), fresh controller, closed loop with
    the reference's own plant (integrate_twist) and growing replay memory"""
    mu = np.array([[2.5, 2.5], [8.5, 2.5]]) + np.array([bounds[0], bounds[2]])
    r = make_oracle(model, mu=mu, lib=RefLib)
    x = np.array([bounds[0] + 5.0, bounds[2] + 7.0, 0.3])
    rec = dict(x=[], ut_before=[], u0=[], ut_after=[], ck=[], mem=[])
    mem = []
    for _ in range(steps):
        rec["x"].append(x.copy())
        rec["ut_before"].append(r.get_ut())
        rec["mem"].append(np.array(mem).reshape(-1, 3))
        u0 = r.control(bounds, x, trace=True)
        rec["u0"].append(u0)
        rec["ut_after"].append(r.get_ut())
        rec["ck"].append(r.last()["ck"])
        x = plant(x[None], u0[None])[0]
        r.add_state_memory(x)
        mem.append(x.copy())
    out = {k: np.array(v) for k, v in rec.items() if k != "mem"}
    out["mem_final"] = np.array(mem)
    out["phik"] = r.get_phik()
    out["bounds"] = np.array(bounds)
    out["mu"] = mu
    out["model"] = np.array(model)
    return out


def batch_case(model, nb, horizon, B, seed, stored=0):
    """configs C2/C4/C5 shapes at a batch the literal reference finishes in seconds"""
    rng = np.random.default_rng(seed)
    steps = int(abs(horizon / 0.1))
    x = random_states(rng, B)
    ut = warm_ut(rng, B, steps, model)
    mem = np.stack([random_states(rng, B) for _ in range(stored)]) if stored else np.zeros((0, B, 3))
    u0, ut_after, ck = [], [], []
    phik = None
    for i in range(B):
        r = make_oracle(model, nb=nb, horizon=horizon, lib=RefLib)
        r.set_ut(ut[i])
        for m in mem:
            r.add_state_memory(m[i])
        u0.append(r.control(BOUNDS_10, x[i], trace=True))
        ut_after.append(r.get_ut())
        ck.append(r.last()["ck"])
        phik = r.get_phik()
    return dict(x=x, ut_before=ut, mem=mem, u0=np.array(u0), ut_after=np.array(ut_after), ck=np.array(ck),
                phik=phik, model=np.array(model), nb=np.array(nb), horizon=np.array(horizon),
                bounds=np.array(BOUNDS_10))


def main():
    assert RefLib.available(), "build oracle/_ref first: make -C oracle ref"
    cases = {
        "c1_cart_10m": closed_loop_case(MODEL_SIMPLE_CART, BOUNDS_10),
        "c1_omni_10m": closed_loop_case(MODEL_OMNI, BOUNDS_10),
        "c1_cart_maze": closed_loop_case(MODEL_SIMPLE_CART, BOUNDS_MAZE),
        "c1_omni_maze": closed_loop_case(MODEL_OMNI, BOUNDS_MAZE),
        "c2_omni_nb10": batch_case(MODEL_OMNI, 10, 5.0, 16, 2),
        "c4_cart_nb20": batch_case(MODEL_SIMPLE_CART, 20, 10.0, 8, 4),
        "c5_omni_nb16_mem100": batch_case(MODEL_OMNI, 16, 5.0, 8, 5, stored=100),
        "omni_nb10_mem10": batch_case(MODEL_OMNI, 10, 5.0, 8, 6, stored=10),
    }
    for name, data in cases.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
