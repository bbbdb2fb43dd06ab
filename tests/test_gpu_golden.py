"""-m gpu: the CUDA path (through the C ABI) against the committed golden
vectors generated from the unmodified reference sources.  Tolerances: see
tests/helpers.py (1e-9 relative, FP64)."""
import glob
import os

import numpy as np
import pytest

from helpers import assert_abs_rel_close, assert_coeff_close, make_gpu

pytestmark = pytest.mark.gpu
# the control() fixtures (make_golden.py); c3_phik_8192 / models_entropy have their own tests
GOLDEN = sorted(f for f in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if os.path.basename(f).startswith(("c1_", "c2_", "c4_", "c5_", "omni_")))


@pytest.mark.parametrize("path", [p for p in GOLDEN if os.path.basename(p).startswith("c1_")], ids=os.path.basename)
def test_closed_loop_cases(path):
    g = np.load(path)
    bounds, model = tuple(g["bounds"]), int(g["model"])
    gpu = make_gpu(model, 1, mu=g["mu"])
    for s in range(len(g["x"])):
        gpu.set_ut(g["ut_before"][s][None])
        u0 = gpu.control(bounds, g["x"][s][None])
        assert_abs_rel_close(u0[0], g["u0"][s], f"step {s} u0")
        assert_abs_rel_close(gpu.get_ut()[0], g["ut_after"][s], f"step {s} ut")
        assert_coeff_close(gpu.get_ck()[0], g["ck"][s], f"step {s} c_k")
        gpu.addStateMemory(g["mem_final"][s][None])
    assert_coeff_close(gpu.get_phik()[0], g["phik"], "phi_k")


@pytest.mark.parametrize("path", [p for p in GOLDEN if not os.path.basename(p).startswith("c1_")],
                         ids=os.path.basename)
def test_batch_cases(path):
    g = np.load(path)
    model, nb, horizon = int(g["model"]), int(g["nb"]), float(g["horizon"])
    B = len(g["x"])
    gpu = make_gpu(model, B, nb=nb, horizon=horizon)
    gpu.set_ut(g["ut_before"])
    for m in g["mem"]:
        gpu.addStateMemory(m)
    u0 = gpu.control(tuple(g["bounds"]), g["x"])
    assert_abs_rel_close(u0, g["u0"], "u0")
    assert_abs_rel_close(gpu.get_ut(), g["ut_after"], "ut")
    ck = gpu.get_ck()
    for i in range(B):
        assert_coeff_close(ck[i], g["ck"][i], f"c_k[{i}]")
    assert_coeff_close(gpu.get_phik()[0], g["phik"], "phi_k")
