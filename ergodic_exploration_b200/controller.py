"""Host-side mirror of the reference's controller API over the C ABI.

``ErgodicControl`` keeps the reference's method names and argument meaning
(ergodic_control.hpp:90-134: control, optTraj, addStateMemory, timeStep,
setTarget, configTarget) for a BATCH of B independent instances living on one
GPU.  numpy arrays go through the ``_host`` entry points (copies + sync, the
end-to-end path); torch CUDA tensors go through the ``_dev`` entry points on
torch's current stream (device-resident path).  Every number is computed by
the CUDA kernels in csrc/ -- this file only marshals pointers.

Array convention: Armadillo column-major 3xN == numpy (N, 3); batched buffers
are (B, N, 3) / (B, 3) / (B, K).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import capi
from .capi import MODEL_CART, MODEL_MECANUM, MODEL_OMNI, MODEL_SIMPLE_CART, EbConfig, ErgodicB200Error, check

try:  # torch is plumbing only (device memory, streams); optional for the host path
    import torch
except Exception:  # pragma: no cover
    torch = None


@dataclass
class Gaussian:
    """target.hpp:56-107 -- 2-D Gaussian with diagonal covariance"""

    mu: Sequence[float]
    sigmas: Sequence[float]


@dataclass
class Target:
    """target.hpp:110-161 -- list of Gaussians (addGaussian / deleteGaussian)"""

    gaussians: list = field(default_factory=list)

    def addGaussian(self, g: Gaussian) -> None:
        self.gaussians.append(g)

    def deleteGaussian(self, idx: int) -> None:
        del self.gaussians[idx]


class _Model:
    """operator() / fdx / fdu / wheels2Twist of the reference's model structs, evaluated by model_eval_kernel
    (csrc/model_kernels.cuh) for one state or a batch (rows)."""

    model_id = -1
    state_space = 3
    params = ()

    def _eval(self, x, u, want):
        lib = capi.load()
        nu = lib.eb_model_controls(self.model_id)
        x = _np_f64(x).reshape(-1, 3)
        u = _np_f64(u).reshape(-1, nu)
        n = x.shape[0]
        if u.shape[0] != n:
            raise ValueError("x and u need the same number of rows")
        f, A, B, vb = np.empty((n, 3)), np.empty((n, 3, 3)), np.empty((n, nu, 3)), np.empty((n, 3))
        par = _np_f64(self.params) if len(self.params) else None
        st = lib.eb_model_eval_host(0, self.model_id, par.ctypes.data if par is not None else None, x.ctypes.data,
                                    u.ctypes.data, n, f.ctypes.data if "f" in want else None,
                                    A.ctypes.data if "A" in want else None, B.ctypes.data if "B" in want else None,
                                    vb.ctypes.data if "vb" in want else None)
        if st == capi.EB_ERR_INVALID_ARGUMENT:
            raise ValueError(lib.eb_last_error().decode())  # std::invalid_argument (cart.hpp:167-170)
        check(st)
        # column-major 3 x 3 / 3 x nu blocks -> numpy (row, col)
        return f, np.transpose(A, (0, 2, 1)), np.transpose(B, (0, 2, 1)), vb

    def __call__(self, x, u):
        f = self._eval(x, u, "f")[0]
        return f[0] if np.ndim(x) == 1 else f

    def fdx(self, x, u):
        A = self._eval(x, u, "A")[1]
        return A[0] if np.ndim(x) == 1 else A

    def fdu(self, x):
        nu = capi.load().eb_model_controls(self.model_id)
        xs = _np_f64(x).reshape(-1, 3)
        B = self._eval(xs, np.zeros((xs.shape[0], nu)), "B")[2]
        return B[0] if np.ndim(x) == 1 else B

    def wheels2Twist(self, u):
        nu = capi.load().eb_model_controls(self.model_id)
        us = _np_f64(u).reshape(-1, nu)
        vb = self._eval(np.zeros((us.shape[0], 3)), us, "vb")[3]
        return vb[0] if np.ndim(u) == 1 else vb


class SimpleCart(_Model):
    """models/cart.hpp:152-206"""

    model_id = MODEL_SIMPLE_CART


class Omni(_Model):
    """models/omni.hpp:164-215"""

    model_id = MODEL_OMNI


class Cart(_Model):
    """models/cart.hpp:60-145 -- 2-wheel differential drive, controls = wheel velocities [uL, uR]"""

    model_id = MODEL_CART

    def __init__(self, wheel_radius: float, wheel_base: float):
        self.wheel_radius, self.wheel_base = float(wheel_radius), float(wheel_base)
        self.params = (self.wheel_radius, self.wheel_base)


class Mecanum(_Model):
    """models/omni.hpp:59-157 -- 4 mecanum wheels, controls = wheel velocities [u0..u3]"""

    model_id = MODEL_MECANUM

    def __init__(self, wheel_radius: float, wheel_base_x: float, wheel_base_y: float):
        self.wheel_radius, self.wheel_base_x, self.wheel_base_y = float(wheel_radius), float(wheel_base_x), float(wheel_base_y)
        self.params = (self.wheel_radius, self.wheel_base_x, self.wheel_base_y)


class RungeKutta:
    """integrator.hpp:60-129 -- the forward problem: solve(model, x0, ut, horizon) -> xt (steps, 3), batched when
    x0 has rows.  Every step is computed by rk4_solve_kernel (one thread per instance)."""

    def __init__(self, dt: float):
        self.dt = float(dt)

    def solve(self, model, x0, ut, horizon: float, device: int = 0):
        lib = capi.load()
        nu = lib.eb_model_controls(model.model_id)
        steps = int(abs(horizon / self.dt))
        single = np.ndim(x0) == 1
        x0 = _np_f64(x0).reshape(-1, 3)
        n = x0.shape[0]
        ut = _np_f64(ut)
        per_instance = ut.ndim == 3
        if ut.shape[-2:] != (steps, nu) and ut.shape[-2:] == (nu, steps):
            ut = np.ascontiguousarray(np.swapaxes(ut, -1, -2))  # Armadillo nu x steps -> (steps, nu) rows
        if ut.shape[-2:] != (steps, nu) or (per_instance and ut.shape[0] != n):
            raise ValueError(f"ut must be (steps={steps}, nu={nu}) or (count, steps, nu)")
        xt = np.empty((n, steps, 3))
        par = _np_f64(model.params) if len(model.params) else None
        st = lib.eb_rk4_solve_host(device, model.model_id, par.ctypes.data if par is not None else None, self.dt,
                                   float(horizon), x0.ctypes.data, ut.ctypes.data, int(per_instance), n, xt.ctypes.data)
        if st == capi.EB_ERR_INVALID_ARGUMENT:
            raise ValueError(lib.eb_last_error().decode())
        check(st)
        return xt[0] if single else xt

    def step(self, model, x, u):
        """one RK4 step (integrator.hpp:176-184), heading NOT wrapped (solve wraps)"""
        raise NotImplementedError("use solve(); the unwrapped single step is not exported")


@dataclass
class GridBounds:
    """The only part of GridMap the controller reads (grid.hpp:214-235)."""

    xmin: float
    xmax: float
    ymin: float
    ymax: float

    def as_tuple(self):
        return (float(self.xmin), float(self.xmax), float(self.ymin), float(self.ymax))


def _is_cuda_tensor(a) -> bool:
    return torch is not None and isinstance(a, torch.Tensor) and a.is_cuda


def _np_f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


class ErgodicControl:
    """Batched drop-in for ErgodicControl<ModelT> (ergodic_control.hpp:72-185)."""

    def __init__(self, model, dt: float, horizon: float, resolution: float, exploration_weight: float,
                 num_basis: int, buffer_size: int, batch_size: int, Rinv, umin, umax, *, batch: int = 1,
                 device: int = 0, seed: int = 0xE16C0D1C, barrier_weight: float = 25.0,
                 barrier_eps: float = 0.05):
        self._lib = capi.load()
        cfg = EbConfig()
        model_id = model if isinstance(model, int) else model.model_id
        self._lib.eb_config_defaults(C.byref(cfg), model_id)
        cfg.batch, cfg.device = int(batch), int(device)
        cfg.dt, cfg.horizon, cfg.resolution = float(dt), float(horizon), float(resolution)
        cfg.expl_weight = float(exploration_weight)
        cfg.num_basis, cfg.buffer_size, cfg.batch_size = int(num_basis), int(buffer_size), int(batch_size)
        R = _np_f64(Rinv, (3, 3))
        for r in range(3):
            for c in range(3):
                cfg.Rinv[r + 3 * c] = R[r, c]  # column-major
        for i in range(3):
            cfg.umin[i] = float(umin[i])
            cfg.umax[i] = float(umax[i])
        cfg.barrier_weight, cfg.barrier_eps, cfg.seed = float(barrier_weight), float(barrier_eps), int(seed)
        h = C.c_void_p()
        st = self._lib.eb_create(C.byref(cfg), C.byref(h))
        if st == capi.EB_ERR_INVALID_ARGUMENT:
            raise ValueError(self._lib.eb_last_error().decode())  # std::invalid_argument in the reference
        check(st)
        self._h = h
        self.cfg = cfg
        self.batch = int(batch)
        self.device = int(device)
        self.steps = self._lib.eb_steps(h)
        self.num_coeff = self._lib.eb_num_coeff(h)
        self.batch_size = int(batch_size)

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.eb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clone(self) -> "ErgodicControl":
        other = object.__new__(ErgodicControl)
        other.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_h"})
        h = C.c_void_p()
        check(self._lib.eb_clone(self._h, C.byref(h)))
        other._h = h
        return other

    def _host_ptr(self, a: np.ndarray) -> int:
        """data pointer of a host array; a control loop passes the same buffers every tick, so the
        last few (array, pointer) pairs are remembered (holding the array keeps its id unique)"""
        cache = self.__dict__.setdefault("_ptrs", {})
        hit = cache.get(id(a))
        if hit is not None and hit[0] is a:
            return hit[1]
        if len(cache) >= 8:
            cache.clear()
        ptr = a.ctypes.data
        cache[id(a)] = (a, ptr)
        return ptr

    def _sync_stream(self):
        if torch is not None and torch.cuda.is_available():
            s = torch.cuda.current_stream(self.device).cuda_stream
            check(self._lib.eb_set_stream(self._h, C.c_void_p(s)))

    # -- reference API ------------------------------------------------------
    def timeStep(self) -> float:
        return self._lib.eb_time_step(self._h)

    def setTarget(self, target) -> None:
        gs = target.gaussians if isinstance(target, Target) else list(target)
        mu = _np_f64([g.mu for g in gs]).reshape(-1)
        sg = _np_f64([g.sigmas for g in gs]).reshape(-1)
        check(self._lib.eb_set_target_gaussians(self._h, len(gs), mu.ctypes.data, sg.ctypes.data))

    def configTarget(self, grid) -> bool:
        b = grid.as_tuple() if hasattr(grid, "as_tuple") else tuple(float(v) for v in grid)
        self._sync_stream()
        rebuilt = C.c_int(0)
        check(self._lib.eb_config_target(self._h, *b, C.byref(rebuilt)))
        return bool(rebuilt.value)

    def reserve_memory(self, count: int) -> None:
        """room for `count` stored states up front (no re-allocation inside the control loop)"""
        check(self._lib.eb_reserve_state_memory(self._h, int(count)))

    def addStateMemory(self, x) -> None:
        if _is_cuda_tensor(x):
            self._sync_stream()
            assert x.dtype == torch.float64 and x.is_contiguous() and x.numel() == 3 * self.batch
            check(self._lib.eb_add_state_memory_dev(self._h, C.c_void_p(x.data_ptr())))
        else:
            x = _np_f64(x).reshape(self.batch, 3)
            check(self._lib.eb_add_state_memory_host(self._h, x.ctypes.data))

    def control(self, grid, x, mem_idx=None, u0=None, metric=None):
        """One control() iteration for all instances.  Returns u0 (B, 3); pass
        ``metric`` (B,) to also receive sum_k lamda_k (c_k - phi_k)^2."""
        if type(grid) is tuple:  # a control loop passes the same bounds tuple every tick
            b = grid
        else:
            b = grid.as_tuple() if hasattr(grid, "as_tuple") else tuple(float(v) for v in grid)
        if _is_cuda_tensor(x):
            self._sync_stream()  # device path: run on torch's current stream
            assert x.dtype == torch.float64 and x.is_contiguous() and x.numel() == 3 * self.batch
            if u0 is None:
                u0 = torch.empty((self.batch, 3), dtype=torch.float64, device=x.device)
            idx_p = C.c_void_p(mem_idx.data_ptr()) if mem_idx is not None else None
            met_p = C.c_void_p(metric.data_ptr()) if metric is not None else None
            st = self._lib.eb_control_dev(self._h, *b, C.c_void_p(x.data_ptr()), idx_p,
                                          C.c_void_p(u0.data_ptr()), met_p)
        else:
            # host path: the call copies / maps the buffers and synchronises by itself
            if not (type(x) is np.ndarray and x.dtype == np.float64 and x.flags.c_contiguous
                    and x.size == 3 * self.batch):
                x = _np_f64(x).reshape(self.batch, 3)
            if u0 is None:
                u0 = np.empty((self.batch, 3))
            idx_p = None
            if mem_idx is not None:
                mem_idx = np.ascontiguousarray(mem_idx, dtype=np.int32)
                idx_p = mem_idx.ctypes.data
            met_p = self._host_ptr(metric) if metric is not None else None
            st = self._lib.eb_control_host(self._h, *b, self._host_ptr(x), idx_p, self._host_ptr(u0), met_p)
        if st != capi.EB_OK:
            if st == capi.EB_ERR_INVALID_ARGUMENT:
                raise ValueError(self._lib.eb_last_error().decode())
            check(st)
        return u0

    def check(self) -> None:
        """Synchronise and surface device-side faults of earlier device-path calls."""
        st = self._lib.eb_check_status(self._h)
        if st == capi.EB_ERR_INVALID_ARGUMENT:
            raise ValueError(self._lib.eb_last_error().decode())
        check(st)

    def optTraj(self, out=None):
        self._sync_stream()
        if _is_cuda_tensor(out):
            check(self._lib.eb_opt_traj_dev(self._h, C.c_void_p(out.data_ptr())))
            return out
        xt = np.empty((self.batch, self.steps, 3))
        st = self._lib.eb_opt_traj_host(self._h, xt.ctypes.data)
        if st == capi.EB_ERR_INVALID_ARGUMENT:
            raise ValueError(self._lib.eb_last_error().decode())
        check(st)
        return xt

    # -- state access (checkpoint / teacher forcing) -------------------------
    def get_ut(self) -> np.ndarray:
        ut = np.empty((self.batch, self.steps, 3))
        check(self._lib.eb_get_ut(self._h, ut.ctypes.data))
        return ut

    def set_ut(self, ut) -> None:
        ut = _np_f64(ut).reshape(self.batch, self.steps, 3)
        check(self._lib.eb_set_ut(self._h, ut.ctypes.data))

    def keep_ck(self, keep: bool) -> None:
        """store (default) or skip the per-instance c_k by-product of control()"""
        check(self._lib.eb_set_keep_ck(self._h, int(bool(keep))))

    def get_ck(self) -> np.ndarray:
        ck = np.empty((self.batch, self.num_coeff))
        check(self._lib.eb_get_ck(self._h, ck.ctypes.data))
        return ck

    def get_phik(self):
        ph = np.empty(self.num_coeff)
        lx, ly = C.c_double(0), C.c_double(0)
        check(self._lib.eb_get_phik(self._h, ph.ctypes.data, C.byref(lx), C.byref(ly)))
        return ph, lx.value, ly.value

    def set_phik(self, phik, lx: float, ly: float) -> None:
        if _is_cuda_tensor(phik):  # e.g. the output of MapTarget.execute / PhikPlan.execute: no host round trip
            self._sync_stream()
            assert phik.dtype == torch.float64 and phik.is_contiguous() and phik.numel() == self.num_coeff
            check(self._lib.eb_set_phik_dev(self._h, C.c_void_p(phik.data_ptr()), float(lx), float(ly)))
            return
        ph = _np_f64(phik).reshape(self.num_coeff)
        check(self._lib.eb_set_phik(self._h, ph.ctypes.data, float(lx), float(ly)))

    def memory_size(self) -> int:
        return int(self._lib.eb_memory_size(self._h))

    def last_mem_idx(self) -> Optional[np.ndarray]:
        n = C.c_int(0)
        idx = np.zeros((self.batch, self.batch_size), dtype=np.int32)
        check(self._lib.eb_get_last_mem_idx(self._h, idx.ctypes.data, C.byref(n)))
        return idx[:, : n.value] if n.value > 0 else None

    def launch_count(self) -> int:
        return int(self._lib.eb_launch_count(self._h))


class PhikPlan:
    """phi_k = (C_y^T Phi C_x) / sum(Phi) for a dense density on the
    configTarget grid (Target::fill normalisation + Basis::spatialCoeff)."""

    def __init__(self, nx: int, ny: int, resolution: float, lx: float, ly: float, nb: int, device: int = 0,
                 algo: int = 0, row_begin: int = 0, ny_total: Optional[int] = None, x_first: float = 0.0,
                 y_first: float = 0.0):
        """ny rows starting at row_begin of an ny_total-row grid (default: the whole grid); x_first / y_first: the
        coordinate of the first sample point (0: the configTarget grid; resolution / 2: cell centres)"""
        self._lib = capi.load()
        h = C.c_void_p()
        ny_total = ny if ny_total is None else ny_total
        st = self._lib.eb_phik_plan_create_ex(device, nx, ny_total, row_begin, ny, float(resolution), float(lx),
                                              float(ly), nb, float(x_first), float(y_first), C.byref(h))
        if st == capi.EB_ERR_INVALID_ARGUMENT:
            raise ValueError(self._lib.eb_last_error().decode())
        check(st)
        self._h, self.nx, self.ny, self.nb, self.device = h, nx, ny, nb, device
        if algo:
            check(self._lib.eb_phik_plan_set_algo(h, algo))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.eb_phik_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def execute(self, phi, phik=None, phi_sum=None):
        if _is_cuda_tensor(phi):
            s = torch.cuda.current_stream(self.device).cuda_stream
            check(self._lib.eb_phik_plan_set_stream(self._h, C.c_void_p(s)))
            assert phi.dtype == torch.float64 and phi.is_contiguous() and phi.numel() == self.nx * self.ny
            if phik is None:
                phik = torch.empty(self.nb * self.nb, dtype=torch.float64, device=phi.device)
            sp = C.c_void_p(phi_sum.data_ptr()) if phi_sum is not None else None
            check(self._lib.eb_phik_execute_dev(self._h, C.c_void_p(phi.data_ptr()), C.c_void_p(phik.data_ptr()), sp))
            return phik
        phi = _np_f64(phi).reshape(self.ny, self.nx)
        out = np.empty(self.nb * self.nb)
        s = C.c_double(0.0)
        check(self._lib.eb_phik_execute_host(self._h, phi.ctypes.data, out.ctypes.data, C.byref(s)))
        self.last_sum = s.value
        return out

    def execute_raw(self, phi, raw=None):
        """un-normalised contraction of this plan's rows: (ld, ld) torch tensor with ld = 32 for nb <= 32, else nb
        rounded up to a multiple of 32; raw[0, 0] = sum(phi)"""
        s = torch.cuda.current_stream(self.device).cuda_stream
        check(self._lib.eb_phik_plan_set_stream(self._h, C.c_void_p(s)))
        assert _is_cuda_tensor(phi) and phi.dtype == torch.float64 and phi.is_contiguous()
        assert phi.numel() == self.nx * self.ny
        if raw is None:
            ld = 32 if self.nb <= 32 else (self.nb + 31) // 32 * 32
            raw = torch.empty((ld, ld), dtype=torch.float64, device=phi.device)
        check(self._lib.eb_phik_execute_raw_dev(self._h, C.c_void_p(phi.data_ptr()), C.c_void_p(raw.data_ptr())))
        return raw

    def launch_count(self) -> int:
        return int(self._lib.eb_phik_launch_count(self._h))

    def fold(self):
        """(mirror fold usable on this grid, measured asymmetry of the cosine table)"""
        f, d = C.c_int(0), C.c_double(0.0)
        check(self._lib.eb_phik_plan_fold(self._h, C.byref(f), C.byref(d)))
        return bool(f.value), d.value


class MapTarget:
    """Map-derived target: int8 occupancy grid -> entropy per cell (numerics.hpp:164-179
    over GridMap::getCell) -> normalised density -> phi_k, all on the device.  One execute per map update."""

    def __init__(self, xsize: int, ysize: int, resolution: float, nb: int, device: int = 0):
        self._lib = capi.load()
        h = C.c_void_p()
        st = self._lib.eb_map_target_create(device, int(xsize), int(ysize), float(resolution), int(nb), C.byref(h))
        if st == capi.EB_ERR_INVALID_ARGUMENT:
            raise ValueError(self._lib.eb_last_error().decode())
        check(st)
        self._h, self.xsize, self.ysize, self.nb, self.device = h, int(xsize), int(ysize), int(nb), int(device)
        self.lx, self.ly = xsize * float(resolution), ysize * float(resolution)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.eb_map_target_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def execute(self, cells, phik=None):
        """cells: (ysize, xsize) int8, torch CUDA tensor (device path, current stream) or numpy (host path)"""
        if _is_cuda_tensor(cells):
            s = torch.cuda.current_stream(self.device).cuda_stream
            check(self._lib.eb_map_target_set_stream(self._h, C.c_void_p(s)))
            assert cells.dtype == torch.int8 and cells.is_contiguous() and cells.numel() == self.xsize * self.ysize
            if phik is None:
                phik = torch.empty(self.nb * self.nb, dtype=torch.float64, device=cells.device)
            check(self._lib.eb_map_target_execute_dev(self._h, C.c_void_p(cells.data_ptr()), C.c_void_p(phik.data_ptr()), None))
            return phik
        cells = np.ascontiguousarray(cells, dtype=np.int8).reshape(self.ysize, self.xsize)
        out = np.empty(self.nb * self.nb)
        s = C.c_double(0.0)
        check(self._lib.eb_map_target_execute_host(self._h, cells.ctypes.data, out.ctypes.data, C.byref(s)))
        self.last_sum = s.value
        return out

    def density(self):
        """the un-normalised entropy density of the last execute, (ysize, xsize) torch view of device memory"""
        from .sharding import _DevView
        ptr = self._lib.eb_map_target_density_dev(self._h)
        return torch.as_tensor(_DevView(ptr, (self.ysize, self.xsize)), device=torch.device("cuda", self.device))

    def launch_count(self) -> int:
        return int(self._lib.eb_map_target_launch_count(self._h))


def l2_gather_peak(device: int = 0) -> float:
    """measured G sectors/s of random 1-byte loads over a 16 MB L2-resident buffer"""
    lib = capi.load()
    v = C.c_double(0)
    check(lib.eb_l2_gather_peak(device, C.byref(v)))
    return v.value


def fp64_peak(device: int = 0):
    """Measured FP64 TFLOP/s of the device: (DFMA loop, DMMA loop)."""
    lib = capi.load()
    a, b = C.c_double(0), C.c_double(0)
    check(lib.eb_fp64_peak(device, C.byref(a), C.byref(b)))
    return a.value, b.value
