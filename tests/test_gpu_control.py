"""-m gpu: parity of the CUDA control() path (through the C ABI) against the
CPU oracle on the same seeded inputs.  Tolerances: see tests/helpers.py."""
import os

import numpy as np
import pytest

from helpers import (BOUNDS_10, BOUNDS_MAZE, MODEL_OMNI, MODEL_SIMPLE_CART, assert_abs_rel_close,
                     assert_angle_close, assert_coeff_close, make_gpu, make_oracle, plant, random_states,
                     warm_ut)

pytestmark = pytest.mark.gpu


def _oracle_step(oracles, bounds, x, mem_idx=None):
    """one control() on every oracle instance; returns u0, c_k, metric, ut"""
    B = len(oracles)
    u0 = np.empty((B, 3))
    ck = np.empty((B, oracles[0].K))
    metric = np.empty(B)
    ut = np.empty((B, oracles[0].steps, 3))
    for i, o in enumerate(oracles):
        u0[i] = o.control(bounds, x[i], mem_idx=None if mem_idx is None else mem_idx[i])
        last = o.last()
        ck[i], metric[i] = last["ck"], last["metric"]
        ut[i] = o.get_ut()
    return u0, ck, metric, ut


def _compare_step(gpu, oracles, bounds, x, mem_idx=None, tag=""):
    metric = np.empty(gpu.batch)
    u0 = gpu.control(bounds, x, mem_idx=mem_idx, metric=metric)
    ou0, ock, ometric, out = _oracle_step(oracles, bounds, x, mem_idx)
    assert_abs_rel_close(u0, ou0, f"{tag} u0")
    assert_abs_rel_close(gpu.get_ut(), out, f"{tag} ut")
    gck = gpu.get_ck()
    for i in range(gpu.batch):
        assert_coeff_close(gck[i], ock[i], f"{tag} c_k[{i}]")
    assert_abs_rel_close(metric, ometric, f"{tag} metric")
    return u0


@pytest.mark.parametrize("model", [MODEL_SIMPLE_CART, MODEL_OMNI])
@pytest.mark.parametrize("bounds", [BOUNDS_10, BOUNDS_MAZE])
def test_c1_single_instance_closed_loop(model, bounds):
    """config C1: fresh controller, then 5 successive steps, teacher-forced
    (both sides are re-synchronised to the oracle's ut_ before every step)."""
    mu = np.array([[2.5, 2.5], [8.5, 2.5]]) + np.array([bounds[0], bounds[2]])
    gpu = make_gpu(model, 1, mu=mu)
    orc = [make_oracle(model, mu=mu)]
    x = np.array([[bounds[0] + 5.0, bounds[2] + 7.0, 0.3]])
    for step in range(6):
        u0 = _compare_step(gpu, orc, bounds, x, tag=f"step {step}")
        gpu.set_ut(orc[0].get_ut()[None])  # teacher forcing (closed loop is chaotic)
        x = plant(x, u0)
    ph, lx, ly = gpu.get_phik()
    assert_coeff_close(ph, orc[0].get_phik(), "phi_k")
    assert abs(lx - (bounds[1] - bounds[0])) < 1e-12 and abs(ly - (bounds[3] - bounds[2])) < 1e-12


@pytest.mark.parametrize("model", [MODEL_SIMPLE_CART, MODEL_OMNI])
def test_free_running_three_steps(model):
    """<= 3 free-running steps stay within tolerance (no teacher forcing)."""
    gpu = make_gpu(model, 1)
    orc = [make_oracle(model)]
    x = np.array([[5.0, 7.0, 0.3]])
    for step in range(3):
        u0 = _compare_step(gpu, orc, BOUNDS_10, x, tag=f"free step {step}")
        x = plant(x, u0)


@pytest.mark.parametrize("model,nb,horizon", [
    (MODEL_OMNI, 10, 5.0),       # C2 shape
    (MODEL_SIMPLE_CART, 20, 10.0),  # C4 shape
    (MODEL_OMNI, 16, 5.0),       # C5 shape
    (MODEL_OMNI, 32, 3.3),
    (MODEL_SIMPLE_CART, 7, 0.7),  # nb padded to the 8-wide instantiation, 7 steps
    (MODEL_OMNI, 11, 6.5),        # nb padded to 12, 65 steps (3 rounds of 32)
    (MODEL_OMNI, 24, 0.2),        # two steps: the minimum the ctor accepts
])
def test_batched_warm_state(model, nb, horizon):
    """random initial states and warm control signals, batch of independent instances"""
    rng = np.random.default_rng(0xE16C0D1C + nb)
    B = 37
    gpu = make_gpu(model, B, nb=nb, horizon=horizon)
    orcs = [make_oracle(model, nb=nb, horizon=horizon) for _ in range(B)]
    x = random_states(rng, B)
    ut = warm_ut(rng, B, gpu.steps, model)
    gpu.set_ut(ut)
    for i, o in enumerate(orcs):
        o.set_ut(ut[i])
    for step in range(2):
        u0 = _compare_step(gpu, orcs, BOUNDS_10, x, tag=f"nb={nb} step {step}")
        gpu.set_ut(np.stack([o.get_ut() for o in orcs]))
        x = plant(x, u0)


@pytest.mark.parametrize("stored", [1, 10, 100, 101, 250])
def test_replay_memory_branches(stored):
    """empty / <= batch_size (all states, insertion order) / > batch_size
    (explicit sample indices) branches of ReplayBuffer::sampleMemory"""
    rng = np.random.default_rng(stored)
    B, model, bs = 5, MODEL_OMNI, 100
    gpu = make_gpu(model, B, batch_size=bs)
    orcs = [make_oracle(model, batch_size=bs) for _ in range(B)]
    for _ in range(stored):
        past = random_states(rng, B)
        gpu.addStateMemory(past)
        for i, o in enumerate(orcs):
            o.add_state_memory(past[i])
    assert gpu.memory_size() == stored
    x = random_states(rng, B)
    ut = warm_ut(rng, B, gpu.steps, model)
    gpu.set_ut(ut)
    for i, o in enumerate(orcs):
        o.set_ut(ut[i])
    idx = rng.integers(0, stored, size=(B, bs)).astype(np.int32) if stored > bs else None
    _compare_step(gpu, orcs, BOUNDS_10, x, mem_idx=idx, tag=f"stored={stored}")


def test_device_sampler_is_replayable():
    """with > batch_size stored states and no indices given, the device draws
    them; feeding the drawn indices to the oracle reproduces the result"""
    rng = np.random.default_rng(7)
    B, model, bs = 4, MODEL_OMNI, 16
    gpu = make_gpu(model, B, batch_size=bs)
    orcs = [make_oracle(model, batch_size=bs) for _ in range(B)]
    for _ in range(40):
        past = random_states(rng, B)
        gpu.addStateMemory(past)
        for i, o in enumerate(orcs):
            o.add_state_memory(past[i])
    x = random_states(rng, B)
    u0 = gpu.control(BOUNDS_10, x)
    idx = gpu.last_mem_idx()
    assert idx is not None and idx.shape == (B, bs) and idx.min() >= 0 and idx.max() < 40
    assert len(np.unique(idx)) > bs // 2  # not degenerate
    ou0, _, _, _ = _oracle_step(orcs, BOUNDS_10, x, idx)
    assert_abs_rel_close(u0, ou0, "device-sampled u0")


@pytest.mark.parametrize("model", [MODEL_SIMPLE_CART, MODEL_OMNI])
def test_opt_traj(model):
    rng = np.random.default_rng(3)
    B = 9
    gpu = make_gpu(model, B, horizon=7.0)
    orcs = [make_oracle(model, horizon=7.0) for _ in range(B)]
    x = random_states(rng, B)
    ut = warm_ut(rng, B, gpu.steps, model)
    gpu.set_ut(ut)
    for i, o in enumerate(orcs):
        o.set_ut(ut[i])
    _compare_step(gpu, orcs, BOUNDS_10, x)
    xt = gpu.optTraj()
    oxt = np.stack([o.opt_traj() for o in orcs])
    assert_abs_rel_close(xt[..., :2], oxt[..., :2], "optTraj xy")
    assert_angle_close(xt[..., 2], oxt[..., 2], "optTraj theta")
    assert np.all(xt[..., 2] >= -np.pi - 1e-12) and np.all(xt[..., 2] < np.pi + 1e-12)


def test_barrier_active_near_walls():
    """states hugging / outside the map edge exercise gradBarrier"""
    B, model = 6, MODEL_OMNI
    gpu = make_gpu(model, B)
    orcs = [make_oracle(model) for _ in range(B)]
    x = np.array([[0.01, 5.0, 0.0], [9.99, 5.0, 3.0], [5.0, 0.02, -1.5], [5.0, 9.98, 1.5],
                  [-0.2, -0.1, 0.7], [10.3, 10.2, -2.0]])
    rng = np.random.default_rng(11)
    ut = warm_ut(rng, B, gpu.steps, model)
    gpu.set_ut(ut)
    for i, o in enumerate(orcs):
        o.set_ut(ut[i])
    _compare_step(gpu, orcs, BOUNDS_10, x, tag="barrier")


def test_simple_cart_rejects_lateral_velocity():
    """SimpleCart throws std::invalid_argument for |u(1)| >= 1e-12 (cart.hpp:167-170)"""
    gpu = make_gpu(MODEL_SIMPLE_CART, 2)
    ut = np.zeros((2, gpu.steps, 3))
    ut[1, 5, 1] = 0.25
    gpu.set_ut(ut)
    with pytest.raises(ValueError, match="y-velocity"):
        gpu.control(BOUNDS_10, np.array([[5.0, 5.0, 0.0], [4.0, 4.0, 1.0]]))


def test_ctor_rejects_single_step_horizon():
    """ergodic_control.hpp:212-216"""
    with pytest.raises(ValueError, match="two steps"):
        make_gpu(MODEL_OMNI, 1, horizon=0.1, dt=0.1)


def test_map_growth_rebuilds_target():
    """configTarget rebuilds phi_k only when the extent changes (:374-377)"""
    gpu = make_gpu(MODEL_OMNI, 1)
    orc = make_oracle(MODEL_OMNI)
    x = np.array([[5.0, 7.0, 0.3]])
    assert gpu.configTarget(BOUNDS_10) is True
    assert gpu.configTarget(BOUNDS_10) is False
    shifted = (1.0, 11.0, -2.0, 8.0)  # same extent, new origin: no rebuild, new map_pos
    assert gpu.configTarget(shifted) is False
    bigger = (0.0, 12.5, 0.0, 11.0)
    for b in (BOUNDS_10, shifted, bigger):
        u0 = gpu.control(b, x)
        ou0 = orc.control(b, x[0])
        assert_abs_rel_close(u0[0], ou0, f"bounds {b}")
        assert_coeff_close(gpu.get_phik()[0], orc.get_phik(), f"phi_k {b}")
        gpu.set_ut(orc.get_ut()[None])


def test_clone_is_deep():
    rng = np.random.default_rng(5)
    gpu = make_gpu(MODEL_OMNI, 3)
    x = random_states(rng, 3)
    gpu.addStateMemory(x)
    gpu.control(BOUNDS_10, x)
    twin = gpu.clone()
    a = gpu.control(BOUNDS_10, x)
    b = twin.control(BOUNDS_10, x)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(gpu.get_ut(), twin.get_ut())


def test_sharding_is_bit_invariant():
    """instances are independent: running a batch as two shards gives
    bit-identical results (multi-GPU logic tested on one GPU)"""
    rng = np.random.default_rng(9)
    B, model = 24, MODEL_OMNI
    x = random_states(rng, B)
    ut = warm_ut(rng, B, 50, model)
    whole = make_gpu(model, B)
    whole.set_ut(ut)
    u_whole = whole.control(BOUNDS_10, x)
    parts = []
    for lo, hi in ((0, 10), (10, 24)):
        g = make_gpu(model, hi - lo)
        g.set_ut(ut[lo:hi])
        parts.append(g.control(BOUNDS_10, x[lo:hi]))
    np.testing.assert_array_equal(u_whole, np.concatenate(parts))


def test_device_path_matches_host_path():
    import torch

    rng = np.random.default_rng(13)
    B, model = 64, MODEL_OMNI
    x = random_states(rng, B)
    ut = warm_ut(rng, B, 50, model)
    a = make_gpu(model, B)
    a.set_ut(ut)
    u_host = a.control(BOUNDS_10, x)
    b = make_gpu(model, B)
    b.set_ut(ut)
    xd = torch.from_numpy(x).cuda()
    md = torch.empty(B, dtype=torch.float64, device="cuda")
    ud = b.control(BOUNDS_10, xd, metric=md)
    b.check()
    np.testing.assert_array_equal(u_host, ud.cpu().numpy())
    assert b.launch_count() >= 1


@pytest.mark.parametrize("nb", [1, 2, 3, 5, 8, 9, 13, 17, 21, 25, 31])
def test_every_basis_count_path(nb):
    """num_basis values between the instantiated tile sizes run on the next
    larger instantiation with the extra coefficients masked"""
    rng = np.random.default_rng(100 + nb)
    B, model = 5, MODEL_OMNI if nb % 2 else MODEL_SIMPLE_CART
    gpu = make_gpu(model, B, nb=nb, horizon=3.0)
    orcs = [make_oracle(model, nb=nb, horizon=3.0) for _ in range(B)]
    x = random_states(rng, B)
    ut = warm_ut(rng, B, gpu.steps, model)
    gpu.set_ut(ut)
    for i, o in enumerate(orcs):
        o.set_ut(ut[i])
    for past in (random_states(rng, B) for _ in range(7)):
        gpu.addStateMemory(past)
        for i, o in enumerate(orcs):
            o.add_state_memory(past[i])
    _compare_step(gpu, orcs, BOUNDS_10, x, tag=f"nb={nb}")


@pytest.mark.parametrize("nb", [33, 40, 64, 100])
def test_wide_basis_counts(nb):
    """num_basis > 32 (the reference accepts any count, basis.cpp:48-77): the CTA-per-instance kernel
    (csrc/solve_kernel_big.cuh) and the phi_k contraction over blocks of 32 orders"""
    rng = np.random.default_rng(300 + nb)
    B, model = 3, MODEL_OMNI if nb % 2 else MODEL_SIMPLE_CART
    gpu = make_gpu(model, B, nb=nb, horizon=3.5)
    orcs = [make_oracle(model, nb=nb, horizon=3.5) for _ in range(B)]
    x = random_states(rng, B)
    ut = warm_ut(rng, B, gpu.steps, model)
    gpu.set_ut(ut)
    for i, o in enumerate(orcs):
        o.set_ut(ut[i])
    for past in (random_states(rng, B) for _ in range(37)):
        gpu.addStateMemory(past)
        for i, o in enumerate(orcs):
            o.add_state_memory(past[i])
    u0 = _compare_step(gpu, orcs, BOUNDS_10, x, tag=f"nb={nb}")
    assert_coeff_close(gpu.get_phik()[0], orcs[0].get_phik(), f"nb={nb} phi_k")
    _compare_step(gpu, orcs, BOUNDS_10, plant(x, u0), tag=f"nb={nb} step 2")


def test_wide_basis_persistent_ctas_and_sampled_memory(monkeypatch):
    """fewer resident CTAs than instances (every CTA walks several instances) with more stored states than
    batch_size (explicit sample indices), two 32-step rounds"""
    monkeypatch.setenv("EB_BIG_ROWS", "2")
    rng = np.random.default_rng(36)
    B, nb, model = 7, 36, MODEL_OMNI
    gpu = make_gpu(model, B, nb=nb, horizon=4.0, batch_size=20)
    orcs = [make_oracle(model, nb=nb, horizon=4.0, batch_size=20) for _ in range(B)]
    x = random_states(rng, B)
    ut = warm_ut(rng, B, gpu.steps, model)
    gpu.set_ut(ut)
    for i, o in enumerate(orcs):
        o.set_ut(ut[i])
    for past in (random_states(rng, B) for _ in range(45)):
        gpu.addStateMemory(past)
        for i, o in enumerate(orcs):
            o.add_state_memory(past[i])
    idx = rng.integers(0, 45, size=(B, 20)).astype(np.int32)
    _compare_step(gpu, orcs, BOUNDS_10, x, mem_idx=idx, tag="nb=36 sampled")
    # the on-device sampler: replay its indices through the oracle
    metric = np.empty(B)
    x2 = random_states(rng, B)
    u0 = gpu.control(BOUNDS_10, x2, metric=metric)
    used = gpu.last_mem_idx()
    ou0, _, ometric, _ = _oracle_step(orcs, BOUNDS_10, x2, used)
    assert_abs_rel_close(u0, ou0, "nb=36 device sampler u0")
    assert_abs_rel_close(metric, ometric, "nb=36 device sampler metric")


def test_wide_basis_host_paths_clone_and_faults():
    """num_basis > 32 through the rest of the API: zero-copy host buffers == pageable buffers == device tensors (bit for
    bit), optTraj from the pose the kernel recorded, a deep clone, c_k switched off, SimpleCart's lateral-velocity fault"""
    import torch

    from ergodic_exploration_b200 import ErgodicB200Error

    rng = np.random.default_rng(64)
    B, nb, model = 6, 40, MODEL_OMNI
    ut = warm_ut(rng, B, 50, model)
    a, b, c = make_gpu(model, B, nb=nb), make_gpu(model, B, nb=nb), make_gpu(model, B, nb=nb)
    for g in (a, b, c):
        g.set_ut(ut)
    c.keep_ck(False)
    x = random_states(rng, B)
    xh = torch.empty((B, 3), dtype=torch.float64).pin_memory()
    uh = torch.empty((B, 3), dtype=torch.float64).pin_memory()
    xh.numpy()[:] = x
    ua = a.control(BOUNDS_10, xh.numpy(), u0=uh.numpy()).copy()  # zero-copy
    ub = b.control(BOUNDS_10, x.copy())                           # pageable
    uc = c.control(BOUNDS_10, torch.from_numpy(x).cuda()).cpu().numpy()  # device tensors, no c_k dump
    c.check()
    np.testing.assert_array_equal(ua, ub)
    np.testing.assert_array_equal(ua, uc)
    np.testing.assert_array_equal(a.optTraj(), b.optTraj())
    o = make_oracle(model, nb=nb)
    o.set_ut(ut[2])
    assert_abs_rel_close(ua[2], o.control(BOUNDS_10, x[2]), "wide row 2")
    oxt = o.opt_traj()
    assert_abs_rel_close(a.optTraj()[2][:, :2], oxt[:, :2], "optTraj xy row 2")
    assert_angle_close(a.optTraj()[2][:, 2], oxt[:, 2], "optTraj theta row 2")
    d = a.clone()
    x2 = plant(x, ua)
    np.testing.assert_array_equal(a.control(BOUNDS_10, x2), d.control(BOUNDS_10, x2))
    cart = make_gpu(MODEL_SIMPLE_CART, 2, nb=nb)
    bad = np.zeros((2, cart.steps, 3))
    bad[1, 3, 1] = 0.25  # cart.hpp:167-170
    cart.set_ut(bad)
    with pytest.raises(ValueError, match="y-velocity"):
        cart.control(BOUNDS_10, random_states(rng, 2))


def test_long_horizon_many_rounds():
    """200 horizon steps = 7 rounds of 32 time-step lanes"""
    rng = np.random.default_rng(77)
    B, model = 3, MODEL_OMNI
    gpu = make_gpu(model, B, nb=6, horizon=20.0)
    orcs = [make_oracle(model, nb=6, horizon=20.0) for _ in range(B)]
    assert gpu.steps == 200
    x = random_states(rng, B)
    ut = warm_ut(rng, B, gpu.steps, model) * 0.3
    gpu.set_ut(ut)
    for i, o in enumerate(orcs):
        o.set_ut(ut[i])
    _compare_step(gpu, orcs, BOUNDS_10, x, tag="N=200")


def test_other_dt_weights_and_limits():
    """non-default dt, exploration weight, control weights and limits"""
    import ergodic_exploration_b200 as eb
    from oracle.pyoracle import Oracle

    rng = np.random.default_rng(21)
    B = 6
    R = np.array([[0.7, 0.1, 0.0], [0.05, 1.3, 0.02], [0.0, 0.3, 2.5]])  # a full (non-diagonal) Rinv
    umin, umax = np.array([-0.4, -0.6, -1.1]), np.array([0.9, 0.3, 0.8])
    mu, sg = [[1.0, 3.0], [4.5, 1.5], [2.0, 2.0]], [[0.4, 0.9], [1.2, 0.3], [0.7, 0.7]]
    bounds = (-2.0, 4.0, 0.5, 4.5)
    gpu = eb.ErgodicControl(eb.Omni(), 0.05, 1.35, 0.07, 3.5, 9, 500, 20, R, umin, umax, batch=B)
    gpu.setTarget([eb.Gaussian(m, s) for m, s in zip(mu, sg)])
    orcs = []
    for _ in range(B):
        o = Oracle.create(MODEL_OMNI, 0.05, 1.35, 0.07, 3.5, 9, 500, 20, R, umin, umax)
        o.set_target(mu, sg)
        orcs.append(o)
    assert gpu.steps == orcs[0].steps == 27
    x = random_states(rng, B, bounds=bounds, margin=0.2)
    ut = rng.uniform(umin, umax, size=(B, gpu.steps, 3))
    gpu.set_ut(ut)
    for i, o in enumerate(orcs):
        o.set_ut(ut[i])
    for step in range(3):
        u0 = _compare_step(gpu, orcs, bounds, x, tag=f"custom step {step}")
        gpu.set_ut(np.stack([o.get_ut() for o in orcs]))
        x = plant(x, u0, dt=0.05)
        gpu.addStateMemory(x)
        for i, o in enumerate(orcs):
            o.add_state_memory(x[i])


def test_memory_buffer_drops_when_full():
    """ReplayBuffer::append silently drops states once buffer_size are stored (buffer.cpp:56-61)"""
    rng = np.random.default_rng(31)
    B = 2
    gpu = make_gpu(MODEL_OMNI, B, buffer_size=5, batch_size=100)
    orcs = [make_oracle(MODEL_OMNI, buffer_size=5, batch_size=100) for _ in range(B)]
    for _ in range(9):
        past = random_states(rng, B)
        gpu.addStateMemory(past)
        for i, o in enumerate(orcs):
            o.add_state_memory(past[i])
    assert gpu.memory_size() == 5 and orcs[0].memory_size() == 5
    _compare_step(gpu, orcs, BOUNDS_10, random_states(rng, B), tag="full buffer")


def test_large_batch_properties():
    """BASELINE-size batch (no oracle at this size): finite, within limits,
    duplicated instances give identical rows, metric >= 0"""
    import torch

    rng = np.random.default_rng(41)
    B, model = 1 << 16, MODEL_OMNI
    gpu = make_gpu(model, B, nb=16)
    x = random_states(rng, B)
    x[B // 2:] = x[: B // 2]  # second half duplicates the first
    ut = warm_ut(rng, B // 2, 50, model)
    gpu.set_ut(np.concatenate([ut, ut]))
    xd = torch.from_numpy(x).cuda()
    md = torch.empty(B, dtype=torch.float64, device="cuda")
    u = gpu.control(BOUNDS_10, xd, metric=md).cpu().numpy()
    gpu.check()
    m = md.cpu().numpy()
    assert np.isfinite(u).all() and np.isfinite(m).all() and (m >= 0).all()
    assert (np.abs(u[:, 0]) <= 1).all() and (np.abs(u[:, 1]) <= 1).all() and (np.abs(u[:, 2]) <= 2).all()
    np.testing.assert_array_equal(u[: B // 2], u[B // 2:])
    np.testing.assert_array_equal(m[: B // 2], m[B // 2:])
    # a sample of rows against the oracle
    for i in rng.integers(0, B // 2, 6):
        o = make_oracle(model, nb=16)
        o.set_ut(ut[i])
        assert_abs_rel_close(u[i], o.control(BOUNDS_10, x[i]), f"row {i}")


def test_zero_copy_host_path_matches_copy_path():
    """eb_control_host with page-locked caller buffers (small batch: the kernel
    reads x / writes u0 and metric in place over PCIe) is bit-identical to the
    same call with pageable buffers (H2D copy, kernel, D2H copy)"""
    import torch

    rng = np.random.default_rng(51)
    B, model = 96, MODEL_OMNI
    ut = warm_ut(rng, B, 50, model)
    a, b = make_gpu(model, B), make_gpu(model, B)
    a.set_ut(ut)
    b.set_ut(ut)
    xh = torch.empty((B, 3), dtype=torch.float64).pin_memory()
    uh = torch.empty((B, 3), dtype=torch.float64).pin_memory()
    mh = torch.empty(B, dtype=torch.float64).pin_memory()
    x = random_states(rng, B)
    for _ in range(3):
        xh.numpy()[:] = x
        ua = a.control(BOUNDS_10, xh.numpy(), u0=uh.numpy(), metric=mh.numpy()).copy()
        mb = np.empty(B)
        ub = b.control(BOUNDS_10, x.copy(), metric=mb)
        np.testing.assert_array_equal(ua, ub)
        np.testing.assert_array_equal(mh.numpy(), mb)
        np.testing.assert_array_equal(a.get_ut(), b.get_ut())
        np.testing.assert_array_equal(a.optTraj(), b.optTraj())  # pose_ recorded by the kernel on both paths
        x = plant(x, ua)
    o = make_oracle(model)
    o.set_ut(ut[5])
    # one row against the oracle as well
    c = make_gpu(model, B)
    c.set_ut(ut)
    x0 = random_states(rng, B)
    xh.numpy()[:] = x0
    u = c.control(BOUNDS_10, xh.numpy(), u0=uh.numpy())
    assert_abs_rel_close(u[5], o.control(BOUNDS_10, x0[5]), "zero-copy row 5")


def test_keep_ck_switch():
    """the K x B c_k dump is optional; switching it off must not change u0"""
    rng = np.random.default_rng(52)
    B, model = 8, MODEL_SIMPLE_CART
    ut = warm_ut(rng, B, 50, model)
    a, b = make_gpu(model, B), make_gpu(model, B)
    a.set_ut(ut)
    b.set_ut(ut)
    b.keep_ck(False)
    x = random_states(rng, B)
    np.testing.assert_array_equal(a.control(BOUNDS_10, x), b.control(BOUNDS_10, x))
    assert a.get_ck().shape == (B, 100)
    with pytest.raises(Exception):
        b.get_ck()


def test_wide_cta_path_single_wave_batch():
    """batches between 4 x #SM and kWideWarps x #SM instances run one wide CTA per SM (28 warps at nb = 10,
    24 at nb = 16, 16 at nb = 20): same numbers as the oracle on sampled instances, and bit-identical to the
    4-warp launch of the same kernel (EB_SOLVE_WIDE only changes the launch shape)"""
    import subprocess
    import sys

    for model, nb, horizon, B in ((MODEL_OMNI, 10, 5.0, 4096), (MODEL_OMNI, 16, 5.0, 3000), (MODEL_SIMPLE_CART, 20, 10.0, 2000)):
        rng = np.random.default_rng(nb)
        gpu = make_gpu(model, B, nb=nb, horizon=horizon)
        x = random_states(rng, B)
        ut = warm_ut(rng, B, gpu.steps, model)
        gpu.set_ut(ut)
        metric = np.empty(B)
        u0 = gpu.control(BOUNDS_10, x, metric=metric)
        ut_new = gpu.get_ut()
        for i in rng.choice(B, 12, replace=False):
            o = make_oracle(model, nb=nb, horizon=horizon)
            o.set_ut(ut[i])
            want = o.control(BOUNDS_10, x[i])
            assert_abs_rel_close(u0[i], want, f"u0 nb={nb} inst {i}")
            assert_abs_rel_close(ut_new[i], o.get_ut(), f"ut_ nb={nb} inst {i}")
            assert_abs_rel_close(metric[i], o.last()["metric"], "metric")
    # launch-shape invariance, in a fresh process with the wide launch switched off
    code = ("import numpy as np, sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from helpers import *\nfrom oracle.pyoracle import MODEL_OMNI\n"
            "rng = np.random.default_rng(10); gpu = make_gpu(MODEL_OMNI, 4096, nb=10)\n"
            "x = random_states(rng, 4096); ut = warm_ut(rng, 4096, gpu.steps, MODEL_OMNI); gpu.set_ut(ut)\n"
            "u0 = gpu.control(BOUNDS_10, x); np.save(sys.argv[1], u0)\n") % (os.path.dirname(os.path.abspath(__file__)),
                                                                          os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import tempfile
    outs = []
    for wide in ("1", "0"):
        with tempfile.NamedTemporaryFile(suffix=".npy") as f:
            subprocess.check_call([sys.executable, "-c", code, f.name], env=dict(os.environ, EB_SOLVE_WIDE=wide))
            outs.append(np.load(f.name))
    assert np.array_equal(outs[0], outs[1])


def test_replay_index_out_of_range_is_reported():
    """buffer.cpp:84,103: memory_.at(i) throws std::out_of_range; here EB_ERR_OUT_OF_RANGE after the step"""
    from ergodic_exploration_b200 import ErgodicB200Error, capi

    rng = np.random.default_rng(3)
    B, bs = 8, 4
    gpu = make_gpu(MODEL_OMNI, B, batch_size=bs)
    for _ in range(bs + 3):
        gpu.addStateMemory(random_states(rng, B))
    x = random_states(rng, B)
    idx = rng.integers(0, bs + 3, size=(B, bs)).astype(np.int32)
    gpu.control(BOUNDS_10, x, mem_idx=idx)  # fine
    idx[2, 1] = bs + 3  # one past the stored states
    with pytest.raises(ErgodicB200Error) as e:
        gpu.control(BOUNDS_10, x, mem_idx=idx)
    assert e.value.status == capi.EB_ERR_OUT_OF_RANGE
    idx[2, 1] = -1
    with pytest.raises(ErgodicB200Error):
        gpu.control(BOUNDS_10, x, mem_idx=idx)
    idx[2, 1] = 0
    gpu.control(BOUNDS_10, x, mem_idx=idx)  # the controller stays usable


def test_horizon_too_long_for_shared_memory_is_refused_at_construction():
    from ergodic_exploration_b200 import ErgodicB200Error, capi

    with pytest.raises(ErgodicB200Error) as e:
        make_gpu(MODEL_OMNI, 4, nb=10, horizon=200.0)  # 2000 steps x 8 fields x 4 warps > 227 KB
    assert e.value.status == capi.EB_ERR_UNSUPPORTED
    make_gpu(MODEL_OMNI, 4, nb=10, horizon=60.0).close()  # 600 steps still fit


def test_nan_guard_reports_non_finite_controls():
    rng = np.random.default_rng(4)
    from ergodic_exploration_b200 import ErgodicB200Error

    gpu = make_gpu(MODEL_OMNI, 4)
    x = random_states(rng, 4)
    x[1, 0] = np.nan
    with pytest.raises(ErgodicB200Error):
        gpu.control(BOUNDS_10, x)
