#!/usr/bin/env python
"""bench.py -- ergodic control solves/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|c5|c3]

One "step" = one batched ErgodicControl::control() iteration (one fused CUDA
kernel) over the workload's instances.  The default workload is BASELINE.json
configs[1] ("c2": Omni, 10x10 basis, 4096 instances, 50-step horizon, shared
two-Gaussian target, warm control signals, no replay memory).  Under torchrun
(N > 1) every rank owns the same number of instances (weak scaling), there is
no data-path collective, and the first twists are gathered with one NCCL
all_gather per step inside the timed region.

Printed JSON (one line, rank 0):
  value      solves/s with inputs resident in HBM (CUDA events on the launching
             stream, one event pair per step, L2 flushed between steps, max over ranks)
  e2e        solves/s through the host-buffer C-ABI call (pinned host x in,
             u0 out, H2D + kernel + D2H + sync inside the timed region)
  roofline   the fused solve kernel against the MEASURED FP64 peak
             (eb_fp64_peak: DFMA / DMMA probes; MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, built from
             the unmodified reference sources) on all host cores, bounded sample

--impl reference times that CPU implementation alone on the same config.
--workload c3 reports the phi_k contraction (grid cells*bases/s) instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: model, nb, horizon, batch per GPU, memory states, description
    "c2": dict(model=1, nb=10, horizon=5.0, batch=4096, mem=0,
               desc="configs[1]: Omni, 10x10 basis, 4096 instances, 50-step horizon"),
    "c2big": dict(model=1, nb=10, horizon=5.0, batch=262144, mem=0,
                  desc="configs[1] shape at 64x the batch: Omni, 10x10 basis, 262144 instances, 50-step horizon"),
    "c4": dict(model=0, nb=20, horizon=10.0, batch=131072, mem=0,
               desc="configs[3] shard: SimpleCart, 20x20 basis, 131072 instances/GPU, 100-step horizon"),
    "c5": dict(model=1, nb=16, horizon=5.0, batch=65536, mem=100,
               desc="configs[4] step: Omni, 16x16 basis, 65536 instances, 50-step horizon, 100 replay states"),
}
BOUNDS = (0.0, 10.0, 0.0, 10.0)
MU = [[2.5, 2.5], [8.5, 2.5]]
SIGMA = [[1.5, 1.5], [1.5, 1.5]]
DT = 0.1


def model_params(model):
    if model == 1:
        return np.diag([1.0, 1.0, 2.0]), np.array([-1.0, -1.0, -2.0]), np.array([1.0, 1.0, 2.0])
    return np.diag([1.0, 0.0, 2.0]), np.array([-1.0, 0.0, -2.0]), np.array([1.0, 0.0, 2.0])


def flops_per_solve(K, N, M):
    """algorithmic FP64 work of one control()"""
    return 2 * K * (N + M) + 4 * K * N + 2 * K + 200 * N


def bytes_per_solve(N, M):
    """ut_ read + write, x, u0, metric, memory gather"""
    return 2 * 24 * N + 24 + 24 + 8 + 24 * M


def synth_inputs(wl, batch, seed):
    """synthetic inputs: x0 ~ U([0.5,9.5]^2 x [-pi,pi)), warm ut_ ~ 0.5 U(umin,umax)"""
    rng = np.random.default_rng(seed)
    _, umin, umax = model_params(wl["model"])
    steps = int(abs(wl["horizon"] / DT))
    x = np.column_stack([rng.uniform(0.5, 9.5, batch), rng.uniform(0.5, 9.5, batch),
                         rng.uniform(-np.pi, np.pi, batch)])
    ut = rng.uniform(umin, umax, size=(batch, steps, 3)) * 0.5
    mem = np.stack([np.column_stack([rng.uniform(0.5, 9.5, batch), rng.uniform(0.5, 9.5, batch),
                                     rng.uniform(-np.pi, np.pi, batch)]) for _ in range(wl["mem"])]) \
        if wl["mem"] else None
    return x, ut, mem


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------
# CPU reference arm
# --------------------------------------------------------------------------
def cpu_reference_rate(wl, sample, steps, warmup, threads=None):
    """Times the reference's CPU control() (oracle/_ref when the compiled
    reference is available, else the C port) on `sample` instances of the
    workload, spread over all host cores, one controller object per instance."""
    import ctypes as C

    from oracle import pyoracle
    from oracle.pyoracle import Oracle, RefLib

    pyoracle.build()
    use_ref = RefLib.available()
    lib = RefLib if use_ref else Oracle
    threads = threads or os.cpu_count() or 1
    x, ut, mem = synth_inputs(wl, sample, seed=0xE16C0D1C + 2)
    R, umin, umax = model_params(wl["model"])
    ctrls = []
    for i in range(sample):
        c = lib.create(wl["model"], DT, wl["horizon"], 0.1, 1.0, wl["nb"], 1000000, 100, R, umin, umax)
        c.set_target(MU, SIGMA)
        c.set_ut(ut[i])
        if mem is not None:
            for m in mem:
                c.add_state_memory(m[i])
        ctrls.append(c)
    handles = (C.c_void_p * sample)(*[c._h for c in ctrls])
    u0 = np.zeros((sample, 3))
    clib = lib.lib()
    chunks = [(lo, min(sample, lo + (sample + threads - 1) // threads))
              for lo in range(0, sample, (sample + threads - 1) // threads)]
    b = [C.c_double(v) for v in BOUNDS]
    dp = C.POINTER(C.c_double)

    def work(lo, hi):
        hs = C.cast(C.byref(handles, lo * C.sizeof(C.c_void_p)), C.POINTER(C.c_void_p))
        xp = x[lo:hi].ctypes.data_as(dp)
        up = u0[lo:hi].ctypes.data_as(dp)
        if use_ref:
            clib.ref_control_many(hs, hi - lo, *b, C.c_double(0.1), xp, up)
        else:
            clib.eo_control_many(hs, hi - lo, *b, xp, up)

    def one_step():
        ts = [threading.Thread(target=work, args=ch) for ch in chunks]
        [t.start() for t in ts]
        [t.join() for t in ts]

    for _ in range(max(1, warmup)):  # first call builds phi_k (excluded)
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    return {
        "value": sample * steps / dt,
        "unit": "solves/s",
        "cores": len(chunks),
        "kind": "reference" if use_ref else "port",
        "sample": f"{sample} instances x {steps} control() steps of the workload, {len(chunks)} host threads, "
                  f"one controller object per instance"
                  + ("" if use_ref else " (C port: compiled reference oracle/_ref not present)"),
        "ms_per_step": dt / steps * 1e3,
    }


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    sample = min(wl["batch"], args.ref_sample)
    r = cpu_reference_rate(wl, sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "ergodic control solves/sec (batched)", "value": r["value"],
        "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "instances_per_step": sample,
                   "note": "reference CPU path (single-threaded per instance), all host cores"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
def run_ours(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import ergodic_exploration_b200 as eb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = wl["batch"]
    R, umin, umax = model_params(wl["model"])
    N = int(abs(wl["horizon"] / DT))
    K = wl["nb"] ** 2
    x, ut, mem = synth_inputs(wl, B, seed=0xE16C0D1C + 2 + rank)
    ctl = eb.ErgodicControl(wl["model"], DT, wl["horizon"], 0.1, 1.0, wl["nb"], 1000000, 100, R, umin, umax,
                            batch=B, device=local_rank)
    ctl.setTarget([eb.Gaussian(m, s) for m, s in zip(MU, SIGMA)])
    ctl.set_ut(ut)
    ctl.keep_ck(False)  # control() returns u0 (+ metric); the K x B c_k dump is a debugging by-product
    if mem is not None:
        for m in mem:
            ctl.addStateMemory(m)
    M = min(wl["mem"], 100)

    xd = torch.from_numpy(x).to(dev)
    u0bufs = [torch.empty((B, 3), dtype=torch.float64, device=dev) for _ in range(2)]
    u0d = u0bufs[0]
    metd = torch.empty(B, dtype=torch.float64, device=dev)
    gathered = torch.empty((world * B, 3), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    state = {"i": 0, "gather_done": None}

    # N > 1: the gather of the first twists is fused into the solve kernel (P2P stores into
    # every rank's gathered buffer over NVLink, csrc/peer_gather.cuh); NCCL is only the fallback
    pg, gather_kind = None, "none (single GPU)"
    if world > 1:
        try:
            from ergodic_exploration_b200.sharding import PeerGather
            pg = PeerGather(ctl)
            gather_kind = "fused into the solve kernel: P2P stores over NVLink peer memory + arrival flags"
        except Exception as exc:  # no peer access between these GPUs
            pg = None
            gather_kind = f"NCCL all_gather per step on a side stream (peer mapping failed: {exc})"

    def step_dev():
        """one control() over this rank's instances.  Fused gather: the kernel publishes its rows
        into every rank's gathered buffer (four rotate; reuse is guarded inside the kernel).
        NCCL fallback: the all_gather of step i runs on a side stream and overlaps step i + 1."""
        if pg is not None:
            pg.control(BOUNDS, xd, metric=metd)
            return
        buf = u0bufs[state["i"] & 1]
        state["i"] += 1
        ctl.control(BOUNDS, xd, u0=buf, metric=metd)
        if world > 1:
            main = torch.cuda.current_stream()
            kdone = torch.cuda.Event()
            kdone.record(main)
            if state["gather_done"] is not None:
                main.wait_event(state["gather_done"])
            with torch.cuda.stream(side):
                side.wait_event(kdone)
                dist.all_gather_into_tensor(gathered, buf)
                state["gather_done"] = torch.cuda.Event()
                state["gather_done"].record(side)

    def head_start(steps):
        """Keeps the host ahead of the device: a spin kernel holds the stream
        while the host enqueues the timed steps, so an event pair brackets the
        kernel itself and not the host's launch latency (a ~30 us kernel is
        shorter than one Python -> ctypes -> cudaLaunch round trip)."""
        torch.cuda._sleep(int(min(steps, 400) * 150e-6 * 1.9e9))

    def drain():
        if pg is not None:
            pg.wait(pg.steps)  # every rank's rows of the last step have arrived here
        elif world > 1 and state["gather_done"] is not None:
            torch.cuda.current_stream().wait_event(state["gather_done"])
            state["gather_done"] = None

    for _ in range(max(3, args.warmup)):
        step_dev()
    drain()
    ctl.check()
    torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctl.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    wall0 = time.perf_counter()
    head_start(args.steps)
    for a, b in ev[:-1]:
        flush.zero_()  # evict the previous step's ut_/x from L2 (outside the event pair)
        a.record()
        step_dev()
        b.record()
    ev[-1][0].record()
    drain()  # the last gather's tail is part of the job
    ev[-1][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = ctl.launch_count() - launches0
    ctl.check()
    t_ms = sum(a.elapsed_time(b) for a, b in ev)

    if pg is not None:
        # the fused gather against a collective: every rank's copy must equal the all_gather of the row blocks
        mine = pg.gathered()[rank * B:(rank + 1) * B].clone()
        ref = torch.empty((world * B, 3), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(ref, mine)
        if not torch.equal(ref, pg.gathered()):
            raise SystemExit("bench.py: fused peer gather disagrees with NCCL all_gather")

    # kernel-only duration for the roofline (single rank / no collective in the pair)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    head_start(args.steps)
    for a, b in kev:
        flush.zero_()
        a.record()
        ctl.control(BOUNDS, xd, u0=u0d, metric=metd)
        b.record()
    torch.cuda.synchronize()
    k_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps

    # end to end through the host-buffer C-ABI call
    xh = torch.from_numpy(x).pin_memory()
    u0h = torch.empty((B, 3), dtype=torch.float64).pin_memory()
    xh_np, u0h_np = xh.numpy(), u0h.numpy()
    for _ in range(3):
        ctl.control(BOUNDS, xh_np, u0=u0h_np)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        ctl.control(BOUNDS, xh_np, u0=u0h_np)  # H2D x, kernel, D2H u0 + status, sync
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - e0
    clk = clocks.stop() if rank == 0 else None

    if pg is not None:
        pg.close()
    if world > 1:
        t = torch.tensor([t_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms, e2e_ms = t.tolist()
        e2e_s = e2e_ms / 1e3
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total = world * B * args.steps
    dfma, dmma = eb.fp64_peak(local_rank)
    peak = max(dfma, dmma)
    F = flops_per_solve(K, N, M)
    achieved = F * B / (k_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "solve_traffic.json")))[args.workload]["bytes"]
    except Exception:
        pass
    line = {
        "metric": "ergodic control solves/sec (batched)",
        "value": total / (t_ms * 1e-3),
        "unit": "solves/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": t_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "instances_per_gpu": B, "num_basis": wl["nb"], "horizon_steps": N,
                   "replay_states": M, "l2": "flushed between timed steps (256 MiB memset outside the event pair)",
                   "timing": "sum of per-step CUDA-event pairs on the launching stream, max over ranks; the host "
                             "enqueues ahead of the device (spin-kernel head start), so a pair brackets the "
                             "step's device work, not host launch latency",
                   "ck_by_product": "off",
                   "parallelism": f"instances sharded over {world} GPUs, no data-path collective; gather of u0: {gather_kind}"
                   if world > 1 else "single GPU"},
        "e2e": {"value": world * B * args.steps / e2e_s, "unit": "solves/s",
                "h2d_bytes_per_step": B * 3 * 8, "d2h_bytes_per_step": B * 3 * 8 + 4,
                "ms_per_step": e2e_s / args.steps * 1e3,
                "path": ("eb_control_host with pinned host buffers: the fused kernel reads x and writes u0 in place over "
                         "PCIe (zero-copy, batch <= 16384), fault flag in mapped memory, stream sync") if B <= 16384 else
                        "eb_control_host: pinned host x -> H2D -> fused kernel -> D2H u0 + fault flag -> sync"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"kernel": "solve_kernel", "bound": "fp64", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "flops_per_solve": F, "kernel_ms": k_ms,
                     "peak_source": f"measured live by eb_fp64_peak (DFMA {dfma:.2f}, DMMA {dmma:.2f} TFLOP/s); "
                                    "MEASURED_PEAKS.json has no FP64 figure",
                     "hbm": {"achieved": bytes_per_solve(N, M) * B / (k_ms * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "bytes_per_solve": bytes_per_solve(N, M),
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
    }
    if world == 1:
        sample = min(B, 4096)
        line["cpu_baseline"] = {k: v for k, v in cpu_reference_rate(wl, sample, 3, 1).items() if k != "ms_per_step"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_loop(args, rank, world, local_rank):
    """configs[4]: the full receding-horizon loop -- per step addStateMemory(x), control(), and the plant
    x <- integrate_twist(x, u0, 0.1) with the angle wrap (the reference's own constant-twist integrator;
    SURVEY section 8d) -- 65536 Omni instances per GPU, 16x16 basis, replay batch 100 drawn by the on-device
    sampler once more than 100 states are stored.  Everything stays on the device; one event pair
    brackets the whole loop.  N > 1: every rank loops over its own instances, the first twists of every
    step are published to all ranks (fused gather) and each rank waits for the complete step."""
    import torch
    import torch.distributed as dist

    import ergodic_exploration_b200 as eb

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS["c5"]
    B, steps, warm = wl["batch"], args.steps, max(3, args.warmup)
    R, umin, umax = model_params(wl["model"])
    N, K = int(abs(wl["horizon"] / DT)), wl["nb"] ** 2
    x, ut, _ = synth_inputs(dict(wl, mem=0), B, seed=0xE16C0D1C + 5 + rank)
    ctl = eb.ErgodicControl(wl["model"], DT, wl["horizon"], 0.1, 1.0, wl["nb"], steps + warm + 8, 100, R, umin, umax,
                            batch=B, device=local_rank)
    ctl.setTarget([eb.Gaussian(m, s) for m, s in zip(MU, SIGMA)])
    ctl.set_ut(ut)
    ctl.keep_ck(False)
    xd = torch.from_numpy(x).to(dev)
    u0d = torch.empty((B, 3), dtype=torch.float64, device=dev)
    metd = torch.empty(B, dtype=torch.float64, device=dev)
    pg = None
    if world > 1:
        from ergodic_exploration_b200.sharding import PeerGather
        pg = PeerGather(ctl)

    def tick():
        ctl.addStateMemory(xd)  # exploration.hpp:209
        if pg is not None:
            step = pg.control(BOUNDS, xd, metric=metd)
            pg.wait(step)
            mine = pg.gathered(step)[rank * B:(rank + 1) * B]
            eb.integrate_twist(xd, mine, DT, out=xd)
        else:
            ctl.control(BOUNDS, xd, u0=u0d, metric=metd)
            eb.integrate_twist(xd, u0d, DT, out=xd)

    for _ in range(warm):
        tick()
    ctl.check()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = ctl.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        tick()
    b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_ms = a.elapsed_time(b)
    launches = ctl.launch_count() - l0 + steps  # + one integrate_twist kernel per step
    ctl.check()
    clk = clocks.stop() if rank == 0 else None
    metric_now = float(metd.mean())
    inside = bool(((xd[:, 0] > -1) & (xd[:, 0] < 11) & (xd[:, 1] > -1) & (xd[:, 1] < 11)).all())
    if pg is not None:
        pg.close()
    if world > 1:
        t = torch.tensor([t_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = t.item()
        dist.destroy_process_group()
    if rank != 0:
        return
    dfma, dmma = eb.fp64_peak(local_rank)
    peak = max(dfma, dmma)
    # replay states per solve: all stored (<= 100) early on, 100 sampled afterwards
    m_avg = sum(min(warm + i + 1, 100) for i in range(steps)) / steps
    F = flops_per_solve(K, N, m_avg)
    line = {
        "metric": "ergodic control solves/sec (batched)", "value": world * B * steps / (t_ms * 1e-3), "unit": "solves/s",
        "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": t_ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"configs[4]: full receding-horizon loop, {steps} control steps, {B} Omni instances per GPU, "
                               "16x16 basis, replay batch 100 (device sampler), plant = integrate_twist",
                   "l2": "closed loop, no flush: per-step working set (2 x 157 MB of ut_) exceeds L2",
                   "timing": "one CUDA-event pair around the whole loop, max over ranks",
                   "mean_ergodic_metric_at_end": metric_now, "robots_inside_map": inside,
                   "parallelism": (f"instances sharded over {world} GPUs; u0 of every step published to all ranks by "
                                   "the solve kernel, each rank waits for the complete step") if world > 1 else "single GPU"},
        "gpu_launches": int(launches), "clocks": clk,
        "e2e": None,
        "roofline": {"kernel": "solve_kernel", "bound": "fp64", "achieved": F * B * steps / (t_ms * 1e-3) / 1e12,
                     "peak": peak, "unit": "TFLOP/s", "frac": F * B * steps / (t_ms * 1e-3) / 1e12 / peak, "traffic": None,
                     "flops_per_solve": F, "note": "whole loop (addStateMemory copy + solve + plant), not the kernel alone",
                     "peak_source": f"measured live by eb_fp64_peak (DFMA {dfma:.2f}, DMMA {dmma:.2f} TFLOP/s)"},
    }
    print(json.dumps(line), flush=True)


def run_phik(args, rank, world, local_rank):
    """secondary metric: phi_k grid cells*bases/sec (configs[2]: 8192^2 grid, 32x32 basis).
    N > 1: the grid is row-sharded (strong scaling of the one contraction), every rank
    contracts its row block and one all_reduce of the raw 32x32 block finishes it."""
    import torch
    import torch.distributed as dist

    import ergodic_exploration_b200 as eb
    from ergodic_exploration_b200.sharding import finish_phik, shard_bounds

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nx = ny = 8192
    nb, res = 32, 0.1
    lx = ly = (nx - 1) * res
    lo, hi = shard_bounds(ny, world, rank)
    g = torch.Generator(device=dev).manual_seed(0xE16C0D1C + 3)
    xs = torch.arange(nx, device=dev, dtype=torch.float64) * res
    ys = torch.arange(lo, hi, device=dev, dtype=torch.float64) * res
    phi = torch.zeros((hi - lo, nx), dtype=torch.float64, device=dev)
    for _ in range(8):  # un-normalised mixture of 8 Gaussians; same stream on every rank
        mu = (0.1 + 0.8 * torch.rand(2, generator=g, device=dev, dtype=torch.float64)) * lx
        sg = (0.02 + 0.08 * torch.rand(2, generator=g, device=dev, dtype=torch.float64)) * lx
        phi += torch.exp(-0.5 * ((xs[None, :] - mu[0]) / sg[0]) ** 2 - 0.5 * ((ys[:, None] - mu[1]) / sg[1]) ** 2)
    algo = int(os.environ.get("EB_PHIK_ALGO", "0"))
    plan = eb.PhikPlan(nx, hi - lo, res, lx, ly, nb, device=local_rank, algo=algo, row_begin=lo, ny_total=ny)
    fold, fold_dev = plan.fold()
    fold = fold and algo != 3
    raw = torch.empty((32, 32), dtype=torch.float64, device=dev)
    out = torch.empty(nb * nb, dtype=torch.float64, device=dev)

    def step():
        if world > 1:
            plan.execute_raw(phi, raw)
            return finish_phik(raw, nb)[0]
        return plan.execute(phi, out)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = plan.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda._sleep(int(min(args.steps, 400) * 100e-6 * 1.9e9))  # host enqueues ahead of the device
    for a, b in ev:
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = plan.launch_count() - l0
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps

    # end to end (N = 1): pageable/pinned host density in, phi_k out, through eb_phik_execute_host
    e2e = None
    if world == 1:
        phih = phi.cpu().pin_memory()
        phin = phih.numpy()
        plan.execute(phin)
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 5))
        for _ in range(reps):
            plan.execute(phin)
        e2e_s = (time.perf_counter() - t0) / reps
        e2e = {"value": nx * ny * nb * nb / e2e_s, "unit": "cell*bases/s", "h2d_bytes_per_step": 8 * nx * ny,
               "d2h_bytes_per_step": 8 * nb * nb + 8, "ms_per_step": e2e_s * 1e3,
               "path": "eb_phik_execute_host: pinned host density -> H2D (512 MiB, PCIe-bound) -> tile kernel -> D2H phi_k"}
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        dist.destroy_process_group()
    if rank != 0:
        return
    dfma, dmma = eb.fp64_peak(local_rank)
    peak64 = max(dfma, dmma)
    flops = 2.0 * nx * ny * nb + 2.0 * ny * nb * nb  # SURVEY §8(d) algorithmic count (unfolded)
    done_flops = flops / 2 if fold else flops       # the fold halves the DMMA work actually issued
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "solve_traffic.json")))["c3"]["bytes"]
    except Exception:
        pass
    sec = ms * 1e-3 * world  # per-GPU seconds of kernel work behind one step (row shards run concurrently)
    hbm = {"achieved": 8.0 * nx * ny / world / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
           "frac": 8.0 * nx * ny / world / (ms * 1e-3) / 1e9 / hbm_peak,
           "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}
    f64 = {"achieved": done_flops / world / (ms * 1e-3) / 1e12, "peak": peak64, "unit": "TFLOP/s",
           "frac": done_flops / world / (ms * 1e-3) / 1e12 / peak64,
           "flops_issued_over_algorithmic": 0.5 if fold else 1.0,
           "peak_source": f"measured live by eb_fp64_peak (DFMA {dfma:.2f}, DMMA {dmma:.2f} TFLOP/s)"}
    del sec
    bound = "hbm" if fold else "fp64"
    main_ = hbm if fold else f64
    line = {
        "metric": "phi_k grid cells*bases/sec", "value": nx * ny * nb * nb / (ms * 1e-3), "unit": "cell*bases/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "configs[2]: phi_k over 8192x8192 Gaussian-mixture grid, 32x32 basis",
                   "l2": "input (512 MiB per step) larger than L2", "mirror_fold": bool(fold),
                   "table_asymmetry": fold_dev,
                   "parallelism": f"rows sharded over {world} GPUs, one all_reduce of the raw 32x32 block" if world > 1
                   else "single GPU"},
        "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"kernel": "phik_dmma_kernel", "bound": bound, "achieved": main_["achieved"], "peak": main_["peak"],
                     "unit": main_["unit"], "frac": main_["frac"], "traffic": traffic, "kernel_ms": ms,
                     "hbm": hbm, "fp64": f64},
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1:
        # CPU beside it: the reference's spatialCoeff arithmetic (2K cosines per cell; its K x G temporary would
        # be 550 TB at this size) as streamed by the C restatement, on a bounded sub-grid, one core
        from oracle import pyoracle
        from oracle.pyoracle import Oracle

        pyoracle.build()
        sub = 384
        phis = phi[:sub, :sub].cpu().numpy()
        t0 = time.perf_counter()
        Oracle.phik_from_grid(phis, res, (sub - 1) * res, (sub - 1) * res, nb)
        dt_cpu = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sub * sub * nb * nb / dt_cpu, "unit": "cell*bases/s", "cores": 1, "kind": "port",
                                "sample": f"{sub}x{sub} corner of the grid, 32x32 basis, Basis::spatialCoeff arithmetic "
                                          f"(basis.cpp:122-133) streamed by oracle/ergodic_oracle.c, single thread"}
    print(json.dumps(line), flush=True)


def run_avoid(args, rank, world, local_rank):
    """widened rows: the collision side of the tick that follows control().
    collide: validate_control of one twist per robot; dwa: DynamicWindow::control (3 x 8 x 5 window,
    2 s rollouts) per robot.  One shared 4000 x 4000 map (16 MB int8, L2-resident), explore_omni.yaml
    radii and limits, 2^18 robots per GPU; the robots are block-partitioned over the ranks with no exchange."""
    import torch
    import torch.distributed as dist

    import ergodic_exploration_b200 as eb

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dwa_mode = args.workload == "dwa"
    rng = np.random.default_rng(0xE16C0D1C + 7 + rank)
    n, res, B = 4000, 0.05, 1 << 18
    mrng = np.random.default_rng(0xE16C0D1C + 8)  # the same map on every rank
    data = np.zeros((n, n), dtype=np.int8)
    data[mrng.random((n, n)) < 0.002] = 100
    data[mrng.random((n, n)) < 0.01] = -1
    colp = (0.7, 1.0, 0.2, 0.8)  # explore_omni.yaml:35-38
    grid = eb.GridMap(-100.0, 100.0, -100.0, 100.0, res, data, device=local_rank)
    col = eb.Collision(*colp)
    x0 = np.column_stack([rng.uniform(-98, 98, B), rng.uniform(-98, 98, B), rng.uniform(-np.pi, np.pi, B)])
    u = np.column_stack([rng.uniform(-1, 1, B), rng.uniform(-1, 1, B), rng.uniform(-2, 2, B)])
    xd, ud = torch.from_numpy(x0).to(dev), torch.from_numpy(u).to(dev)
    dwa_cfg = (0.1, 2.0, 0.2, 2.5, 2.5, 1.0, 1.0, -1.0, 1.0, -1.0, 2.0, -2.0)  # explore_omni.yaml:15-28,65-70
    dwa = eb.DynamicWindow(col, *dwa_cfg, 3, 8, 5)
    vref = torch.zeros_like(ud)

    def step():
        if dwa_mode:
            return dwa.control(grid, xd, ud, vref=vref)[0]
        return eb.validate_control(col, grid, xd, ud, 0.1, 0.5)

    def step_host(xh, uh):
        if dwa_mode:
            return dwa.control(grid, xh, uh, vref=np.zeros_like(uh))[0]
        return eb.validate_control(col, grid, xh, uh, 0.1, 0.5)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(max(3, args.warmup)):
        out = step()
    torch.cuda.synchronize()
    free_frac = float(out.float().mean())
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = grid.launch_count()
    steps = min(args.steps, 100)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.zero_()  # the map and the poses leave L2 between steps
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = grid.launch_count() - l0
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    # the pre-dilated map is built once per map update; time one rebuild + step beside the steady state
    build_ms = None
    if dwa_mode:
        grid.update(data)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        torch.cuda.synchronize()
        build_ms = a.elapsed_time(b) - ms
    # end to end: host poses / twists in, flags out
    xh, uh = x0.copy(), u.copy()
    step_host(xh, uh)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        step_host(xh, uh)
    e2e_s = (time.perf_counter() - t0) / reps
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t.tolist()
        dist.destroy_process_group()
    if rank != 0:
        return
    # algorithmic probes of one full (collision-free) check: the reference's circle walk, pruned radii
    def circle_cells(r):
        x, y, err, cnt = -r, 0, 2 - 2 * r, 0
        while x < 0:
            cnt += 4
            rr = err
            if rr <= y:
                y += 1
                err += 2 * y + 1
            if rr > x or err > y:
                x += 1
                err += 2 * x + 1
        return cnt
    r_bnd, r_col, r_max = int(colp[0] / res), int((colp[0] + colp[2]) / res), int(colp[1] / res)
    probes_pose = sum(circle_cells(r) for r in range(r_bnd, min(r_max, r_col) + 1))
    probes_ref = sum(circle_cells(r) for r in range(r_bnd, r_max + 1))
    rollouts = 120 if dwa_mode else 1
    nsteps = 20 if dwa_mode else 5
    unit = "DWA decisions/s" if dwa_mode else "validated twists/s"
    line = {
        "metric": ("DynamicWindow::control decisions/sec (batched)" if dwa_mode else "validate_control twists/sec (batched)"),
        "value": world * B / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8/f64",
        "data": "synthetic",
        "config": {"workload": ("DynamicWindow 3x8x5 window, 2 s rollouts" if dwa_mode else "validate_control, 0.5 s rollout")
                   + f", {B} robots per GPU on one 4000x4000 map @ 0.05 m, explore_omni.yaml radii",
                   "collision_free_fraction": free_frac, "l2": "flushed between timed steps",
                   "dilated_map": ("pose checks are single lookups into a pre-dilated map, rebuilt per map update: "
                                   f"{build_ms:.2f} ms for this map (outside the timed steps)") if dwa_mode else
                                  "not used at this pose count (circle walks)",
                   "parallelism": f"robots block-partitioned over {world} GPU(s), no exchange"},
        "e2e": {"value": world * B / e2e_s, "unit": unit, "h2d_bytes_per_step": 2 * B * 24,
                "d2h_bytes_per_step": B * (4 + (24 if dwa_mode else 0)), "ms_per_step": e2e_s * 1e3,
                "path": "host poses + twists -> H2D -> kernel -> D2H flags" + (" + twists" if dwa_mode else "")},
        "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"kernel": "dwa_control_kernel" if dwa_mode else "validate_control_kernel",
                     "bound": "gather latency / issue (int8 probes, map L2-resident); no closed-form peak",
                     "achieved": B * rollouts * nsteps * probes_pose / (ms * 1e-3) / 1e9, "peak": None,
                     "unit": "G cell probes/s (upper bound: early exits probe less)", "frac": None, "traffic": None,
                     "probes_per_pose": probes_pose, "probes_per_pose_reference_unpruned": probes_ref},
    }
    if world == 1:
        from oracle import pyoracle
        from oracle.pyoracle import Oracle, RefLib

        pyoracle.build()
        lib = RefLib if RefLib.available() else Oracle
        sample = 2048 if dwa_mode else 200000
        sub = slice(0, sample)
        t0 = time.perf_counter()
        if dwa_mode:
            lib.dwa_control(data, res, -100.0, -100.0, colp, dwa_cfg, (3, 8, 5), x0[sub], u[sub], vref=np.zeros((sample, 3)))
        else:
            lib.validate_control(data, res, -100.0, -100.0, colp, x0[sub], u[sub], 0.1, 0.5)
        dt_cpu = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sample / dt_cpu, "unit": unit, "cores": 1,
                                "kind": "reference" if lib is RefLib else "port",
                                "sample": f"{sample} robots of the same workload, the reference's own "
                                          f"{'DynamicWindow::control' if dwa_mode else 'validate_control'} on one core "
                                          "(includes one GridMap construction)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-sample", type=int, default=4096, help="instances per step of the CPU reference arm")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c3", "c5loop", "collide", "dwa"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "c3":
        return run_phik(args, rank, world, local_rank)
    if args.workload in ("collide", "dwa"):
        return run_avoid(args, rank, world, local_rank)
    if args.workload == "c5loop":
        return run_loop(args, rank, world, local_rank)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl, rank, world)
    run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
