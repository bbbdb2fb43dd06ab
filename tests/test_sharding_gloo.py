"""CPU, world_size 2, gloo: the host-side sharding / gather logic of the
multi-GPU path (ergodic_exploration_b200/sharding.py).  Compute is stood in by
the CPU oracle (allowed in tests): the point is that block partition +
all_gather / all_reduce reproduce the unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ergodic_exploration_b200.sharding import all_gather_rows, finish_phik, shard_bounds, shard_sizes


def test_shard_bounds_partition():
    for total in (0, 1, 7, 8, 4096, 1 << 20, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import BOUNDS_10, MODEL_OMNI, make_oracle, random_states, warm_ut
    from oracle.pyoracle import Oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- control(): block partition, local solve, all_gather of u0 -----
        rng = np.random.default_rng(123)  # same stream on every rank: the full problem
        x = random_states(rng, total)
        ut = warm_ut(rng, total, 20, MODEL_OMNI)
        lo, hi = shard_bounds(total, world, rank)
        u_local = np.empty((hi - lo, 3))
        for i in range(lo, hi):
            o = make_oracle(MODEL_OMNI, nb=6, horizon=2.0)
            o.set_ut(ut[i])
            u_local[i - lo] = o.control(BOUNDS_10, x[i])
        u_all = all_gather_rows(torch.from_numpy(u_local), total).numpy()
        # ---- phi_k: row blocks, local raw contraction, all_reduce ----------
        nx, ny, nb, res = 40, 23, 6, 0.1
        phi = np.random.default_rng(5).random((ny, nx))
        lx, ly = (nx - 1) * res, (ny - 1) * res
        rlo, rhi = shard_bounds(ny, world, rank)
        ys, xs = np.cumsum(np.full(ny, res)) - res, np.cumsum(np.full(nx, res)) - res
        cy = np.cos(np.arange(nb)[None, :] * (np.pi / ly) * ys[rlo:rhi, None])
        cx = np.cos(np.arange(nb)[None, :] * (np.pi / lx) * xs[:, None])
        raw = np.zeros((32, 32))
        raw[:nb, :nb] = cy.T @ phi[rlo:rhi] @ cx
        phik, total_phi = finish_phik(torch.from_numpy(raw), nb)
        if rank == 0:
            want_u = np.empty((total, 3))
            for i in range(total):
                o = make_oracle(MODEL_OMNI, nb=6, horizon=2.0)
                o.set_ut(ut[i])
                want_u[i] = o.control(BOUNDS_10, x[i])
            want_phik, want_sum = Oracle.phik_from_grid(phi, res, lx, ly, nb)
            q.put((np.array_equal(u_all, want_u), float(np.max(np.abs(phik.numpy() - want_phik))),
                   abs(float(total_phi) - want_sum)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 11])
def test_world2_gather_matches_unsharded(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    [p.start() for p in procs]
    same_u, phik_err, sum_err = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert same_u, "gathered u0 differs from the unsharded run (must be bit-identical)"
    assert phik_err < 1e-13 and sum_err < 1e-10
