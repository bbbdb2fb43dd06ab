"""Host-side mirror of the reference's occupancy-grid collision API over the C ABI.

``Collision`` (collision.hpp:85-165), ``GridMap`` (grid.hpp:80-260, the parts the
checks read), ``validate_control`` (numerics.hpp:312-330) and ``DynamicWindow``
(dynamic_window.hpp:60-172) keep the
reference's names and argument meaning for a BATCH of poses / candidate twists
sharing one map.  numpy arrays go through the ``_host`` entry points, torch
CUDA tensors through the ``_dev`` ones.  Everything is computed by
csrc/collision_kernels.cuh; this file only marshals pointers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import EbCollision, EbDwa, check

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_cuda(a) -> bool:
    return torch is not None and isinstance(a, torch.Tensor) and a.is_cuda


class Collision:
    """2-D collision detector parameters (collision.hpp:95-96).  Raises ValueError
    where the reference constructor throws std::invalid_argument (collision.cpp:52-63)."""

    def __init__(self, boundary_radius: float, search_radius: float, obstacle_threshold: float,
                 occupied_threshold: float):
        if search_radius < boundary_radius:
            raise ValueError("Search radius must be at least the same size as the boundary radius")
        if occupied_threshold > 100.0 or occupied_threshold < 0.0:
            raise ValueError("Occupied threshold must be between 0 and 100")
        self.cfg = EbCollision(float(boundary_radius), float(search_radius), float(obstacle_threshold),
                               float(occupied_threshold))

    def totalPadding(self) -> float:
        return self.cfg.boundary_radius + self.cfg.obstacle_threshold

    def collisionCheck(self, grid: "GridMap", pose):
        """collision.cpp:126-143 for a batch of poses (B, 3) -> int32 (B,), 1 = collision"""
        return grid._check(self, pose)


class GridMap:
    """Occupancy grid resident on the GPU (grid.hpp:94-95: bounds, resolution, int8 cells
    in row-major order, i = y row / j = x column)."""

    def __init__(self, xmin: float, xmax: float, ymin: float, ymax: float, resolution: float, grid_data,
                 device: int = 0):
        self._lib = capi.load()
        data = np.ascontiguousarray(grid_data, dtype=np.int8)
        xsize = int(round((xmax - xmin) / resolution))  # axis_length grid.hpp:61-64
        ysize = int(round((ymax - ymin) / resolution))
        if xsize * ysize != data.size:
            raise ValueError("Grid data size does not match the grid size")  # grid.cpp:57-60
        self.xsize, self.ysize, self.resolution = xsize, ysize, float(resolution)
        self.xmin, self.xmax, self.ymin, self.ymax = float(xmin), float(xmax), float(ymin), float(ymax)
        self.device = device
        h = C.c_void_p()
        check(self._lib.eb_grid_create(device, data.ctypes.data, xsize, ysize, float(resolution), float(xmin),
                                       float(ymin), C.byref(h)))
        self._h = h

    def as_tuple(self):
        return (self.xmin, self.xmax, self.ymin, self.ymax)

    def update(self, grid_data) -> None:
        data = np.ascontiguousarray(grid_data, dtype=np.int8)
        if data.size != self.xsize * self.ysize:
            raise ValueError("Grid data size does not match the grid size")
        check(self._lib.eb_grid_update(self._h, data.ctypes.data))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.eb_grid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self) -> int:
        return int(self._lib.eb_grid_launch_count(self._h))

    def dilation(self, mode: int) -> None:
        """0 = automatic, 1 = always walk the circles, 2 = always use the pre-dilated map"""
        check(self._lib.eb_grid_set_dilation(self._h, int(mode)))

    def _stream(self):
        if torch is not None and torch.cuda.is_available():
            s = torch.cuda.current_stream(self.device).cuda_stream
            check(self._lib.eb_grid_set_stream(self._h, C.c_void_p(s)))

    def _check(self, collision: Collision, pose):
        self._stream()
        if _is_cuda(pose):
            assert pose.dtype == torch.float64 and pose.is_contiguous() and pose.numel() % 3 == 0
            n = pose.numel() // 3
            out = torch.empty(n, dtype=torch.int32, device=pose.device)
            check(self._lib.eb_collision_check_dev(self._h, C.byref(collision.cfg), C.c_void_p(pose.data_ptr()), n,
                                                   C.c_void_p(out.data_ptr())))
            return out
        pose = np.ascontiguousarray(pose, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pose), dtype=np.int32)
        check(self._lib.eb_collision_check_host(self._h, C.byref(collision.cfg), pose.ctypes.data, len(pose),
                                                out.ctypes.data))
        return out

    def _validate(self, collision: Collision, x0, u, dt: float, horizon: float):
        self._stream()
        if _is_cuda(x0):
            assert _is_cuda(u) and x0.dtype == torch.float64 and u.dtype == torch.float64
            assert x0.is_contiguous() and u.is_contiguous() and x0.numel() == u.numel()
            n = x0.numel() // 3
            out = torch.empty(n, dtype=torch.int32, device=x0.device)
            check(self._lib.eb_validate_control_dev(self._h, C.byref(collision.cfg), C.c_void_p(x0.data_ptr()),
                                                    C.c_void_p(u.data_ptr()), n, float(dt), float(horizon),
                                                    C.c_void_p(out.data_ptr())))
            return out
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, 3)
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, 3)
        if x0.shape != u.shape:
            raise ValueError("x0 and u must have the same shape")
        out = np.empty(len(x0), dtype=np.int32)
        check(self._lib.eb_validate_control_host(self._h, C.byref(collision.cfg), x0.ctypes.data, u.ctypes.data,
                                                 len(x0), float(dt), float(horizon), out.ctypes.data))
        return out


def validate_control(collision: Collision, grid: GridMap, x0, u, dt: float, horizon: float):
    """numerics.hpp:312-330 for a batch: (B, 3) start poses and twists -> int32 (B,),
    1 = the twist held for |horizon / dt| steps stays collision free"""
    return grid._validate(collision, x0, u, dt, horizon)


class DynamicWindow:
    """Dynamic window approach for a batch of robots sharing one map (dynamic_window.hpp:60-65;
    same constructor argument order).  ``control`` mirrors both reference overloads
    (dynamic_window.cpp:93-139 with a reference twist, :141-187 with a reference trajectory)
    and returns (found (B,), u_opt (B, 3)) -- numpy in, numpy out; torch CUDA in, torch out."""

    def __init__(self, collision: Collision, dt: float, horizon: float, acc_dt: float, acc_lim_x: float,
                 acc_lim_y: float, acc_lim_th: float, max_vel_x: float, min_vel_x: float, max_vel_y: float,
                 min_vel_y: float, max_rot_vel: float, min_rot_vel: float, vx_samples: int, vy_samples: int,
                 vth_samples: int):
        self._lib = capi.load()
        self.collision = collision
        self.cfg = EbDwa(float(dt), float(horizon), float(acc_dt), float(acc_lim_x), float(acc_lim_y),
                         float(acc_lim_th), float(max_vel_x), float(min_vel_x), float(max_vel_y), float(min_vel_y),
                         float(max_rot_vel), float(min_rot_vel), int(vx_samples), int(vy_samples), int(vth_samples))

    def timeStep(self) -> float:
        return self.cfg.dt

    def horizon(self) -> float:
        return self.cfg.horizon

    def steps(self) -> int:
        return int(abs(self.cfg.horizon / self.cfg.dt))

    def control(self, grid: GridMap, x0, vb, vref=None, xt_ref=None, dt_ref: float = 0.0, min_cost=None, out=None):
        """vref: (B, 3) reference twists, or xt_ref: a reference trajectory (ncols, 3) shared by the batch
        or (B, ncols, 3) per instance (e.g. ErgodicControl.optTraj()) with its time step dt_ref.
        Host path: ``out=(found int32 (B,), u (B, 3))`` lets a control loop reuse (page-locked) result buffers;
        with every buffer page-locked, large batches are copied and computed in overlapping slices."""
        if (vref is None) == (xt_ref is None):
            raise ValueError("give either vref or xt_ref")
        grid._stream()
        col, cfg = C.byref(self.collision.cfg), C.byref(self.cfg)
        if _is_cuda(x0):
            n = x0.numel() // 3
            found = torch.empty(n, dtype=torch.int32, device=x0.device)
            u = torch.empty((n, 3), dtype=torch.float64, device=x0.device)
            ref = vref if vref is not None else xt_ref
            for t in (x0, vb, ref):
                assert _is_cuda(t) and t.dtype == torch.float64 and t.is_contiguous()
            mc = C.c_void_p(min_cost.data_ptr()) if min_cost is not None else None
            P = lambda t: C.c_void_p(t.data_ptr())
            if vref is not None:
                check(self._lib.eb_dwa_control_twist_dev(grid._h, col, cfg, P(x0), P(vb), P(vref), n, P(found), P(u), mc))
            else:
                per = 1 if xt_ref.dim() == 3 else 0
                ncols = xt_ref.shape[-2]
                check(self._lib.eb_dwa_control_traj_dev(grid._h, col, cfg, P(x0), P(vb), P(xt_ref), ncols, per,
                                                        float(dt_ref), n, P(found), P(u), mc))
            return found, u
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, 3)
        vb = np.ascontiguousarray(vb, dtype=np.float64).reshape(-1, 3)
        n = len(x0)
        if out is not None:
            found, u = out
            assert found.dtype == np.int32 and found.size == n and u.dtype == np.float64 and u.size == 3 * n
            assert found.flags.c_contiguous and u.flags.c_contiguous
        else:
            found, u = np.empty(n, dtype=np.int32), np.empty((n, 3))
        mc = min_cost.ctypes.data if min_cost is not None else None
        if vref is not None:
            vref = np.ascontiguousarray(vref, dtype=np.float64).reshape(-1, 3)
            st = self._lib.eb_dwa_control_twist_host(grid._h, col, cfg, x0.ctypes.data, vb.ctypes.data, vref.ctypes.data,
                                                     n, found.ctypes.data, u.ctypes.data, mc)
        else:
            xt = np.ascontiguousarray(xt_ref, dtype=np.float64)
            per = 1 if xt.ndim == 3 else 0
            st = self._lib.eb_dwa_control_traj_host(grid._h, col, cfg, x0.ctypes.data, vb.ctypes.data, xt.ctypes.data,
                                                    xt.shape[-2], per, float(dt_ref), n, found.ctypes.data,
                                                    u.ctypes.data, mc)
        if st == capi.EB_ERR_INVALID_ARGUMENT:
            raise ValueError(self._lib.eb_last_error().decode())
        check(st)
        return found, u


def integrate_twist(x, u, dt: float, out=None):
    """numerics.hpp:273-298 + the angle wrap :77-89 for (B, 3) torch CUDA poses and twists, on the
    current stream; ``out`` may be ``x`` itself"""
    assert _is_cuda(x) and _is_cuda(u) and x.dtype == torch.float64 and u.dtype == torch.float64
    assert x.is_contiguous() and u.is_contiguous() and x.numel() == u.numel()
    if out is None:
        out = torch.empty_like(x)
    dev = x.device.index or 0
    s = torch.cuda.current_stream(dev).cuda_stream
    check(capi.load().eb_integrate_twist_dev(dev, C.c_void_p(x.data_ptr()), C.c_void_p(u.data_ptr()), float(dt),
                                             x.numel() // 3, C.c_void_p(out.data_ptr()), C.c_void_p(s)))
    return out
