"""The header-only C++ drop-in adapter (include/ergodic_exploration_b200/):
compiled with g++ against the test-only Armadillo stand-in and linked to the
C-ABI library.  CPU: it builds, the host-side pieces match the reference's
known answers, and it refuses to run without a GPU.  -m gpu: a node-main-like
closed loop through ErgodicControl<ModelT>::control() matches the oracle."""
import os
import subprocess

import numpy as np
import pytest

from helpers import (BOUNDS_10, MODEL_OMNI, MODEL_SIMPLE_CART, assert_abs_rel_close, assert_angle_close,
                     assert_coeff_close, make_oracle)
from oracle.pyoracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "adapter_demo")


@pytest.fixture(scope="module")
def demo():
    import __graft_entry__ as g
    from ergodic_exploration_b200 import capi

    if not os.path.exists(capi.LIB_PATH):
        g.build()
    libdir = os.path.join(ROOT, "ergodic_exploration_b200")
    cmd = ["g++", "-std=c++20", "-O1", "-Wall", "-Wextra", "-Wno-unused-parameter", "-DERGODIC_B200_WITH_ROS",
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "shim"),
           os.path.join(ROOT, "tests", "cpp", "adapter_demo.cpp"), "-o", EXE, "-L" + libdir, "-lergodic_b200",
           "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "warning" not in r.stderr, r.stderr
    return EXE


def test_adapter_builds_and_host_checks(demo):
    r = subprocess.run([demo, "cpu"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr


def _parse(out):
    recs, cur = {}, []
    for line in out.splitlines():
        t = line.split()
        if not t or t[0] in ("OK", "steps"):
            continue
        recs.setdefault(t[0], []).append(np.array([float(v) for v in t[1:]]))
    return recs


@pytest.mark.gpu
@pytest.mark.parametrize("model", [MODEL_SIMPLE_CART, MODEL_OMNI])
def test_adapter_closed_loop_matches_oracle(demo, model):
    r = subprocess.run([demo, "gpu", str(model)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    rec = _parse(r.stdout)
    o = make_oracle(model)
    for s in range(4):
        x = rec["x"][s]
        o.add_state_memory(x)
        u0 = o.control(BOUNDS_10, x)
        assert_abs_rel_close(rec["u0"][s], u0, f"step {s} u0")
        assert_abs_rel_close(rec["ut"][s].reshape(-1, 3), o.get_ut(), f"step {s} ut")
        traj, otraj = rec["traj"][s].reshape(-1, 3), o.opt_traj()
        assert_abs_rel_close(traj[:, :2], otraj[:, :2], f"step {s} optTraj")
        assert_angle_close(traj[:, 2], otraj[:, 2], f"step {s} optTraj theta")
        o.set_ut(rec["ut"][s].reshape(-1, 3))  # teacher forcing
    assert_coeff_close(rec["phik"][0], o.get_phik(), "phi_k")
    # Basis / Target public methods
    pts = np.array([[1.0 + 1.5 * i, 9.0 - 1.25 * i] for i in range(6)])
    assert_coeff_close(rec["fk"][0], Oracle.fourier_basis(10.0, 10.0, 10, [1.25, 7.5]), "fourierBasis")
    assert_coeff_close(rec["dfk"][0], Oracle.grad_fourier_basis(10.0, 10.0, 10, [1.25, 7.5]).reshape(-1), "grad")
    assert_coeff_close(rec["ck"][0], Oracle.traj_coeff(10.0, 10.0, 10, pts), "trajCoeff")
    vals = Oracle.target_fill([[2.5, 2.5], [8.5, 2.5]], [[1.5, 1.5], [1.5, 1.5]], [0.0, 0.0], pts)
    assert_coeff_close(rec["fill"][0], vals, "Target::fill")
    assert_coeff_close(rec["sc"][0], Oracle.spatial_coeff(10.0, 10.0, 10, vals, pts), "spatialCoeff")
    # RungeKutta adapter: test/test_integrator.cpp:44-73 on the GPU, and a Mecanum rollout against the oracle
    rk = rec["rkcart"][0].reshape(4, 3)
    for i in range(4):
        assert abs(rk[i, 0] - 0.01 * (i + 1)) <= 4 * np.spacing(0.01 * (i + 1)) and rk[i, 1] == 0.0 and rk[i, 2] == 0.0
    um = np.array([[3.0 * np.sin(1.0 + r + 2.0 * c) for r in range(4)] for c in range(6)])
    want = Oracle.rk4_forward_mecanum(0.05, 0.3, 0.2, 0.05, 0.31, [0.4, -0.2, 1.1], um)
    assert_abs_rel_close(rec["rkmec"][0].reshape(6, 3), want, "Mecanum RK4 through the adapter")
    # batched face
    xb = rec["xb"][0].reshape(3, 3)
    for i in range(3):
        ob = make_oracle(model, buffer_size=1000)
        assert_abs_rel_close(rec["ub"][0].reshape(3, 3)[i], ob.control(BOUNDS_10, xb[i]), f"batched u0[{i}]")
        assert_abs_rel_close(rec["metric"][0][i], ob.last()["metric"], f"batched metric[{i}]")
    # collision checks (include/ergodic_exploration_b200/collision.hpp): the demo's map, rebuilt here
    data = np.zeros((90, 120), dtype=np.int8)
    data[40, 20:100] = 100
    data[10, 10], data[70, 60], data[5, 100] = 100, 77, -1
    col = (0.2, 1.0, 0.05, 0.65)
    poses, twists = rec["cpose"][0].reshape(-1, 3), rec["ctwist"][0].reshape(-1, 3)
    np.testing.assert_array_equal(rec["chit"][0].astype(int), Oracle.collision_check(data, 0.05, -1.0, 0.5, col, poses))
    np.testing.assert_array_equal(rec["cvalid"][0].astype(int),
                                  Oracle.validate_control(data, 0.05, -1.0, 0.5, col, poses, twists, 0.1, 1.0))
    assert 0 < rec["cvalid"][0].sum() < len(poses)
    dwa_cfg = (0.1, 2.0, 0.2, 2.5, 2.5, 1.0, 1.0, -1.0, 1.0, -1.0, 2.0, -2.0)
    fo, uo, _ = Oracle.dwa_control(data, 0.05, -1.0, 0.5, col, dwa_cfg, (3, 8, 5), poses, twists, vref=np.zeros_like(twists))
    np.testing.assert_array_equal(rec["dfound"][0].astype(int), fo)
    np.testing.assert_array_equal(rec["dtwist"][0].reshape(-1, 3), uo)
    ft, ut, _ = Oracle.dwa_control(data, 0.05, -1.0, 0.5, col, dwa_cfg, (3, 8, 5), poses[9:10], twists[9:10], xt_ref=poses,
                                   dt_ref=0.1)
    assert int(rec["dtraj"][0][0]) == ft[0]
    np.testing.assert_array_equal(rec["dtraju"][0], ut[0])
