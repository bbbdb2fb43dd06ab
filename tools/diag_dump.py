#!/usr/bin/env python
"""debug aid: per-step intermediates of instance 0 (EB_DEBUG_DUMP build) against the oracle"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
DBG = "/tmp/libergodic_b200_dump.so"
EXTRA = os.environ.get("EB_EXTRA_FLAGS", "").split()
subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                       "-DEB_DEBUG_DUMP", *EXTRA, "-shared", "-o", DBG,
                       os.path.join(ROOT, "ergodic_exploration_b200", "csrc", "ergodic_b200.cu"), "-lcudart"])
os.environ["EB_LIB_PATH"] = DBG
import numpy as np  # noqa: E402
from helpers import BOUNDS_10, make_gpu, make_oracle, random_states, warm_ut  # noqa: E402
from ergodic_exploration_b200 import capi  # noqa: E402

np.set_printoptions(precision=3, linewidth=220, suppress=False)
model, nb, horizon = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 10, 5.0
rng = np.random.default_rng(7)
B, steps = 2, 50
gpu = make_gpu(model, B, nb=nb, horizon=horizon)
o = make_oracle(model, nb=nb, horizon=horizon)
ut = warm_ut(rng, B, steps, model)
gpu.set_ut(ut)
o.set_ut(ut[0])
x = random_states(rng, B)
gpu.control(BOUNDS_10, x)
o.control(BOUNDS_10, x[0])
last = o.last()
buf = np.zeros((steps, 16))
assert capi.load().eb_debug_dump(buf.ctypes.data_as(C.c_void_p), steps) == 0
print("xf err", np.abs(buf[:, 8] - last["xtf"][:, 0]).max(), "yf err", np.abs(buf[:, 9] - last["xtf"][:, 1]).max())
lx = 10.0
print("ca err", np.abs(buf[:, 2] - np.cos(np.pi * buf[:, 8] / lx)).max(), "sa err", np.abs(buf[:, 3] - np.sin(np.pi * buf[:, 8] / lx)).max())
print("cb err", np.abs(buf[:, 4] - np.cos(np.pi * buf[:, 9] / lx)).max(), "sb err", np.abs(buf[:, 5] - np.sin(np.pi * buf[:, 9] / lx)).max())
print("ce err", np.abs(buf[:, 6] - np.cos(last["xtf"][:, 2])).max(), "se err", np.abs(buf[:, 7] - np.sin(last["xtf"][:, 2])).max())
print("xf err per step", np.abs(buf[:, 8] - last["xtf"][:, 0]))
print("ca err per step", np.abs(buf[:, 2] - np.cos(np.pi * buf[:, 8] / lx)))
print("ex err per step", np.abs(buf[:, 0] - last["edx"][:, 0]))
print("ey err per step", np.abs(buf[:, 1] - last["edx"][:, 1]))
ck = gpu.get_ck()[0]
print("ck err", np.abs(ck - last["ck"]).max())
