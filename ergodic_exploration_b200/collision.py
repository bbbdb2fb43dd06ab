"""Host-side mirror of the reference's occupancy-grid collision API over the C ABI.

``Collision`` (collision.hpp:85-165), ``GridMap`` (grid.hpp:80-260, the parts the
checks read) and ``validate_control`` (numerics.hpp:312-330) keep the
reference's names and argument meaning for a BATCH of poses / candidate twists
sharing one map.  numpy arrays go through the ``_host`` entry points, torch
CUDA tensors through the ``_dev`` ones.  Everything is computed by
csrc/collision_kernels.cuh; this file only marshals pointers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import EbCollision, check

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
