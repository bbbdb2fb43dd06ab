"""CPU: the C-ABI library loads and exports every symbol include/ergodic_b200.h
declares; without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ergodic_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    from ergodic_exploration_b200 import capi

    if not os.path.exists(capi.LIB_PATH):
        g.build()
    return capi.load()


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(eb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_typed(lib):
    from ergodic_exploration_b200 import capi

    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in capi.SIGNATURES, f"{s} declared in the header but not typed in capi.SIGNATURES"
    assert sorted(capi.SIGNATURES) == syms  # and nothing undeclared is bound


def test_abi_version_and_defaults(lib):
    from ergodic_exploration_b200 import capi

    assert lib.eb_abi_version() == 1
    cfg = capi.EbConfig()
    lib.eb_config_defaults(C.byref(cfg), capi.MODEL_OMNI)
    assert (cfg.dt, cfg.horizon, cfg.resolution, cfg.num_basis, cfg.batch_size) == (0.1, 5.0, 0.1, 10, 100)
    assert list(cfg.Rinv) == [1, 0, 0, 0, 1, 0, 0, 0, 2] and list(cfg.umax) == [1, 1, 2]
    assert (cfg.barrier_weight, cfg.barrier_eps) == (25.0, 0.05)
    lib.eb_config_defaults(C.byref(cfg), capi.MODEL_SIMPLE_CART)
    assert list(cfg.Rinv) == [1, 0, 0, 0, 0, 0, 0, 0, 2] and list(cfg.umin) == [-1, 0, -2]


def test_argument_validation_needs_no_gpu(lib):
    from ergodic_exploration_b200 import capi

    cfg = capi.EbConfig()
    lib.eb_config_defaults(C.byref(cfg), capi.MODEL_OMNI)
    h = C.c_void_p()
    cfg.horizon = 0.1  # one step: the reference throws std::invalid_argument (ergodic_control.hpp:212-216)
    assert lib.eb_create(C.byref(cfg), C.byref(h)) == capi.EB_ERR_INVALID_ARGUMENT
    assert b"two steps" in lib.eb_last_error()
    cfg.horizon, cfg.num_basis = 5.0, 129  # EB_MAX_NUM_BASIS = 128
    assert lib.eb_create(C.byref(cfg), C.byref(h)) == capi.EB_ERR_INVALID_ARGUMENT
    cfg.num_basis, cfg.model = 10, 7
    assert lib.eb_create(C.byref(cfg), C.byref(h)) == capi.EB_ERR_INVALID_ARGUMENT
    import torch
    if not torch.cuda.is_available():
        # 33..128 pass validation (the CTA-per-instance kernel): without a device the answer is "no device", not "bad argument"
        cfg.num_basis, cfg.model = 64, capi.MODEL_OMNI
        assert lib.eb_create(C.byref(cfg), C.byref(h)) == capi.EB_ERR_NO_DEVICE
    assert h.value is None
    assert lib.eb_control_host(None, 0, 1, 0, 1, None, None, None, None) == capi.EB_ERR_INVALID_ARGUMENT
    # round-2 entry points
    assert lib.eb_map_target_create(0, 0, 10, 0.05, 8, C.byref(h)) == capi.EB_ERR_INVALID_ARGUMENT
    assert lib.eb_map_target_create(0, 10, 10, 0.05, 129, C.byref(h)) == capi.EB_ERR_INVALID_ARGUMENT
    assert lib.eb_rk4_solve_host(0, 7, None, 0.1, 1.0, None, None, 0, 1, None) == capi.EB_ERR_INVALID_ARGUMENT
    assert lib.eb_model_eval_host(0, 2, None, None, None, 1, None, None, None, None) == capi.EB_ERR_INVALID_ARGUMENT  # Cart needs params
    assert [lib.eb_model_controls(m) for m in range(5)] == [3, 3, 2, 4, 0]
    assert lib.eb_set_phik_dev(None, None, 1.0, 1.0) == capi.EB_ERR_INVALID_ARGUMENT
    assert lib.eb_gather_fuse_min_batch() > 0


def test_new_entry_points_have_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ergodic_exploration_b200 as eb

    with pytest.raises(eb.ErgodicB200Error) as e:
        eb.MapTarget(64, 64, 0.05, 8)
    assert e.value.status == eb.EB_ERR_NO_DEVICE
    with pytest.raises(eb.ErgodicB200Error):
        eb.RungeKutta(0.1).solve(eb.Cart(0.1, 2.0), [0.0, 0.0, 0.0], np.ones((4, 2)), 0.4)
    with pytest.raises(eb.ErgodicB200Error):
        eb.Omni()([0.0, 0.0, 0.0], [1.0, 0.0, 0.0])


def test_no_cpu_fallback(lib):
    """without a CUDA device the product refuses to run instead of computing on the host"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ergodic_exploration_b200 as eb

    assert lib.eb_device_count() == 0
    with pytest.raises(eb.ErgodicB200Error) as e:
        eb.ErgodicControl(eb.Omni(), 0.1, 5.0, 0.1, 1.0, 10, 1000, 100, np.eye(3), [-1] * 3, [1] * 3)
    assert e.value.status == eb.EB_ERR_NO_DEVICE
    with pytest.raises(eb.ErgodicB200Error):
        eb.PhikPlan(16, 16, 0.1, 1.5, 1.5, 4)


def test_product_never_touches_the_oracle():
    """only tests/, smoke() and bench.py's CPU legs may use oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ergodic_exploration_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "capi.py" and False, f"{f} mentions the oracle"
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            assert "pyoracle" not in open(os.path.join(dirpath, f)).read()
