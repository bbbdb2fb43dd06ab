"""-m gpu: the fused gather of the first twists (csrc/peer_gather.cuh).  On one GPU the
peer group has a single member (the kernel stores into its own gathered buffer and raises
its own flag); with two or more GPUs a two-process run checks the rows that arrive over
NVLink against an NCCL all_gather."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import BOUNDS_10, MODEL_OMNI, make_gpu, plant, random_states, warm_ut

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fuse_min_batch", ["0", "1000000"])  # publish from inside the kernel / from the side stream
def test_single_rank_group_matches_plain_control(fuse_min_batch, monkeypatch):
    import torch

    monkeypatch.setenv("EB_GATHER_FUSE_MIN_BATCH", fuse_min_batch)

    from ergodic_exploration_b200.sharding import PeerGather

    rng = np.random.default_rng(61)
    B = 200
    ut = warm_ut(rng, B, 50, MODEL_OMNI)
    a, b = make_gpu(MODEL_OMNI, B), make_gpu(MODEL_OMNI, B)
    a.set_ut(ut)
    b.set_ut(ut)
    pg = PeerGather(a)
    x = random_states(rng, B)
    for step in range(1, 5):
        xd = torch.from_numpy(x).cuda()
        md = torch.empty(B, dtype=torch.float64, device="cuda")
        assert pg.control(BOUNDS_10, xd, metric=md) == step
        pg.wait(step)
        got = pg.gathered(step).cpu().numpy()
        mb = np.empty(B)
        want = b.control(BOUNDS_10, x, metric=mb)
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(md.cpu().numpy(), mb)
        x = plant(x, want)
    np.testing.assert_array_equal(a.get_ut(), b.get_ut())
    pg.close()


WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from helpers import BOUNDS_10, MODEL_OMNI, make_gpu, random_states, warm_ut
from ergodic_exploration_b200.sharding import PeerGather
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
B = 512
rng = np.random.default_rng(100 + rank)
ctl = make_gpu(MODEL_OMNI, B, device=rank)
ctl.set_ut(warm_ut(rng, B, 50, MODEL_OMNI))
pg = PeerGather(ctl)
x = torch.from_numpy(random_states(rng, B)).cuda()
for step in range(1, 14):
    if step % 3 == 0:   # the one-call form: solve + gather + wait on the controller's stream
        assert pg.control_wait(BOUNDS_10, x) == step
    else:
        pg.control(BOUNDS_10, x)
        pg.wait(step)
    torch.cuda.synchronize()
    g = pg.gathered(step)
    ref = torch.empty_like(g)
    dist.all_gather_into_tensor(ref, g[rank * B:(rank + 1) * B].clone())
    assert torch.equal(ref, g), f"rank {rank} step {step}: gathered rows differ from all_gather"
    assert float(g.abs().sum()) > 0
pg.close()
dist.destroy_process_group()
print("PEER_OK", rank)
"""


@pytest.mark.parametrize("fuse_min_batch", ["0", "1000000"])
def test_two_ranks_over_nvlink(tmp_path, fuse_min_batch):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT],
                       capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, EB_GATHER_FUSE_MIN_BATCH=fuse_min_batch))
    assert r.returncode == 0 and r.stdout.count("PEER_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


PHIK_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import ergodic_exploration_b200 as eb
from ergodic_exploration_b200.sharding import PhikAllReduce, finish_phik, shard_bounds
from oracle.pyoracle import Oracle
for nx, ny, nb in ((512, 300, 16), (1030, 257, 20)):
    res = 0.1
    lx, ly = (nx - 1) * res, (ny - 1) * res
    rng = np.random.default_rng(nx + ny)   # the same density on every rank
    phi = rng.random((ny, nx))
    lo, hi = shard_bounds(ny, world, rank)
    plan = eb.PhikPlan(nx, hi - lo, res, lx, ly, nb, device=rank, row_begin=lo, ny_total=ny)
    par = PhikAllReduce(plan)
    mine = torch.from_numpy(np.ascontiguousarray(phi[lo:hi])).cuda()
    want, total = Oracle.phik_from_grid(phi, res, lx, ly, nb)
    for step in range(5):   # several steps: both receive buffers, flags beyond 1
        tot = torch.empty(1, dtype=torch.float64, device="cuda")
        got = par.execute(mine, phi_sum=tot)
        torch.cuda.synchronize()
        err = float(np.max(np.abs(got.cpu().numpy() - want)) / np.max(np.abs(want)))
        assert err <= 1e-9, f"rank {rank} step {step}: fused all-reduce phi_k error {err:.3e}"
        assert abs(float(tot) - total) <= 1e-9 * total
        # every rank holds the same bits
        allg = torch.empty((world, nb * nb), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allg, got)
        assert torch.equal(allg[0], allg[rank])
    # and the NCCL route of the same plan agrees
    raw = plan.execute_raw(mine)
    ref = finish_phik(raw, nb)[0]
    assert float((ref - got).abs().max()) <= 1e-12
    par.close()
dist.destroy_process_group()
print("PHIK_PEER_OK", rank)
"""


def test_fused_phik_allreduce_two_ranks(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker_phik.py"
    script.write_text(PHIK_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29535", str(script), ROOT],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.count("PHIK_PEER_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


def test_fused_phik_allreduce_single_rank_group():
    """a group of one: the kernel takes the single-GPU path, same numbers as the plain plan"""
    import torch

    from ergodic_exploration_b200 import PhikPlan
    from ergodic_exploration_b200.sharding import PhikAllReduce

    rng = np.random.default_rng(9)
    phi = torch.from_numpy(rng.random((200, 256))).cuda()
    plan = PhikPlan(256, 200, 0.1, 25.5, 19.9, 12, algo=4)  # the collective always runs the TMA tile kernel
    par = PhikAllReduce(plan)
    a = par.execute(phi)
    b = plan.execute(phi)
    assert torch.equal(a, b)
    par.close()
