"""ctypes loaders for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``Oracle``  : oracle/libergodic_oracle.so, the plain-C restatement of the
  reference hot path (oracle/ergodic_oracle.c).
* ``RefLib``  : oracle/_ref/libergodic_ref.so, the UNMODIFIED reference
  sources compiled against the test shim (oracle/ref_driver.cpp).  Present in
  the build container (and on the GPU box as a prebuilt file that travelled
  with the snapshot); ``RefLib.available()`` says whether it can be loaded.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(ergodic_exploration_b200/) must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libergodic_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libergodic_ref.so")

MODEL_SIMPLE_CART = 0
MODEL_OMNI = 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    """contiguous float64 view + pointer"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def build(force: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is present, the
    compiled reference (make is a no-op for up-to-date targets)."""
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "ergodic_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/include/ergodic_exploration"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


class _Controller:
    """Shared Python face of eo_controller / RefController."""

    def __init__(self, lib, prefix, handle, nb, steps, is_ref, map_res):
        self._lib, self._p, self._h = lib, prefix, handle
        self.nb, self.K, self.steps = nb, nb * nb, steps
        self._is_ref, self._map_res = is_ref, map_res

    def _f(self, name):
        return getattr(self._lib, self._p + name)

    def close(self):
        if self._h:
            self._f("destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_target(self, mu, sigma):
        mu, pm = _d(mu)
        sigma, ps = _d(sigma)
        self._f("set_target")(self._h, C.c_int(mu.size // 2), pm, ps)

    def add_state_memory(self, x):
        x, px = _d(x)
        self._f("add_state_memory")(self._h, px)

    def memory_size(self):
        return int(self._f("memory_size")(self._h))

    def get_ut(self):
        ut = np.zeros((self.steps, 3))
        self._f("get_ut")(self._h, ut.ctypes.data_as(_dp))
        return ut

    def set_ut(self, ut):
        ut, p = _d(ut)
        assert ut.size == 3 * self.steps
        self._f("set_ut")(self._h, p)

    def get_phik(self):
        ph = np.zeros(self.K)
        if self._is_ref:
            self._f("get_phik")(self._h, ph.ctypes.data_as(_dp), None)
        else:
            self._f("get_phik")(self._h, ph.ctypes.data_as(_dp))
        return ph

    def set_phik(self, phik, lx, ly):
        """phik_ and the basis extent set directly (restatement only): the configTarget of a control() whose map
        extent equals (lx, ly) keeps them (ergodic_control.hpp:374-377)"""
        assert not self._is_ref
        ph, pp = _d(phik)
        assert ph.size == self.K
        self._f("set_phik")(self._h, pp, C.c_double(lx), C.c_double(ly))

    def opt_traj(self):
        xt = np.zeros((self.steps, 3))
        n = self._f("opt_traj")(self._h, xt.ctypes.data_as(_dp))
        if n < 0:
            raise ValueError("SimpleCart: invalid twist y-velocity must be 0")
        return xt

    def control(self, bounds, x, mem_idx=None, trace=False):
        """bounds = (xmin, xmax, ymin, ymax).  Returns u0 (3,)."""
        x, px = _d(x)
        u0 = np.zeros(3)
        b = [C.c_double(float(v)) for v in bounds]
        if self._is_ref:
            fn = self._f("control_trace" if trace else "control")
            rc = fn(self._h, *b, C.c_double(self._map_res), px, u0.ctypes.data_as(_dp))
        else:
            idx = None
            if mem_idx is not None:
                idx_arr = np.ascontiguousarray(mem_idx, dtype=np.int32)
                idx = idx_arr.ctypes.data_as(_ip)
            rc = self._f("control")(self._h, *b, px, idx, u0.ctypes.data_as(_dp))
        if rc != 0:
            raise ValueError("control() failed (SimpleCart guard or missing sample indices)")
        return u0

    def last(self):
        """dict(ck, edx, bdx, rhot, xtf[, metric]) of the last control()/trace."""
        ck = np.zeros(self.K)
        edx, bdx, rhot, xtf = (np.zeros((self.steps, 3)) for _ in range(4))
        ptr = lambda a: a.ctypes.data_as(_dp)
        if self._is_ref:
            self._f("get_last")(self._h, ptr(ck), ptr(edx), ptr(bdx), ptr(rhot), ptr(xtf))
            return dict(ck=ck, edx=edx, bdx=bdx, rhot=rhot, xtf=xtf)
        m = C.c_double(0.0)
        self._f("get_last")(self._h, ptr(ck), C.byref(m), ptr(edx), ptr(bdx), ptr(rhot), ptr(xtf))
        return dict(ck=ck, edx=edx, bdx=bdx, rhot=rhot, xtf=xtf, metric=m.value)


class Oracle:
    """The plain-C restatement."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(ORACLE_SO):
                build()
            lib = C.CDLL(ORACLE_SO)
            lib.eo_normalize_angle_pi.restype = C.c_double
            lib.eo_normalize_angle_pi.argtypes = [C.c_double]
            lib.eo_target_evaluate.restype = C.c_double
            lib.eo_create.restype = C.c_void_p
            lib.eo_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                      C.c_longlong, C.c_int, _dp, _dp, _dp]
            for n in ("destroy", "set_target", "add_state_memory", "get_ut", "set_ut", "get_phik",
                      "get_last", "control", "opt_traj", "memory_size", "steps", "config_target",
                      "set_phik"):
                getattr(lib, "eo_" + n).argtypes = None
            lib.eo_destroy.argtypes = [C.c_void_p]
            lib.eo_set_target.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
            lib.eo_add_state_memory.argtypes = [C.c_void_p, _dp]
            lib.eo_get_ut.argtypes = [C.c_void_p, _dp]
            lib.eo_set_ut.argtypes = [C.c_void_p, _dp]
            lib.eo_get_phik.argtypes = [C.c_void_p, _dp]
            lib.eo_set_phik.argtypes = [C.c_void_p, _dp, C.c_double, C.c_double]
            lib.eo_get_last.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]
            lib.eo_control.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _ip, _dp]
            lib.eo_config_target.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
            lib.eo_opt_traj.argtypes = [C.c_void_p, _dp]
            lib.eo_memory_size.argtypes = [C.c_void_p]
            lib.eo_memory_size.restype = C.c_longlong
            lib.eo_steps.argtypes = [C.c_void_p]
            lib.eo_control_many.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_double, C.c_double,
                                            C.c_double, C.c_double, _dp, _dp]
            lib.eo_rk4_forward.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp, _dp]
            lib.eo_rk4_forward_cart.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp]
            lib.eo_phik_from_grid.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                              C.c_int, _dp, _dp]
            lib.eo_integrate_twist.argtypes = [_dp, _dp, C.c_double, _dp]
            cls._lib = lib
        return cls._lib

    @classmethod
    def create(cls, model, dt, horizon, resolution, expl_weight, num_basis, buffer_size, batch_size,
               Rinv, umin, umax):
        lib = cls.lib()
        Rinv, pR = _d(np.asarray(Rinv, dtype=np.float64).T)  # column-major
        umin, pmin = _d(umin)
        umax, pmax = _d(umax)
        h = lib.eo_create(model, dt, horizon, resolution, expl_weight, num_basis, buffer_size, batch_size,
                          pR, pmin, pmax)
        if not h:
            raise ValueError("Need at least two steps in forward simulation")
        steps = lib.eo_steps(C.c_void_p(h))
        return _Controller(lib, "eo_", C.c_void_p(h), num_basis, steps, False, None)

    # ---- stateless helpers -------------------------------------------------
    @classmethod
    def model_f(cls, model, x, u):
        x, px = _d(x); u, pu = _d(u); out = np.zeros(3)
        rc = cls.lib().eo_model_f(C.c_int(model), px, pu, out.ctypes.data_as(_dp))
        if rc:
            raise ValueError("Invalid twist y-velocity must be 0.")
        return out

    @classmethod
    def model_fdx(cls, model, x, u):
        x, px = _d(x); u, pu = _d(u); A = np.zeros(9)
        cls.lib().eo_model_fdx(C.c_int(model), px, pu, A.ctypes.data_as(_dp))
        return A.reshape(3, 3).T

    @classmethod
    def model_fdu(cls, model, x):
        x, px = _d(x); B = np.zeros(9)
        cls.lib().eo_model_fdu(C.c_int(model), px, B.ctypes.data_as(_dp))
        return B.reshape(3, 3).T

    @classmethod
    def cart(cls, r, b, x, u):
        lib = cls.lib(); x, px = _d(x); u, pu = _d(u)
        f, A, B, tw = np.zeros(3), np.zeros(9), np.zeros(6), np.zeros(3)
        P = lambda a: a.ctypes.data_as(_dp)
        rr, bb = C.c_double(r), C.c_double(b)
        lib.eo_cart_f(rr, bb, px, pu, P(f)); lib.eo_cart_fdx(rr, bb, px, pu, P(A))
        lib.eo_cart_fdu(rr, bb, px, P(B)); lib.eo_cart_wheels2twist(rr, bb, pu, P(tw))
        return f, A.reshape(3, 3).T, B.reshape(2, 3).T, tw

    @classmethod
    def mecanum(cls, r, bx, by, x, u):
        lib = cls.lib(); x, px = _d(x); u, pu = _d(u)
        f, A, B, tw = np.zeros(3), np.zeros(9), np.zeros(12), np.zeros(3)
        P = lambda a: a.ctypes.data_as(_dp)
        a3 = (C.c_double(r), C.c_double(bx), C.c_double(by))
        lib.eo_mecanum_f(*a3, px, pu, P(f)); lib.eo_mecanum_fdx(*a3, px, pu, P(A))
        lib.eo_mecanum_fdu(*a3, px, P(B)); lib.eo_mecanum_wheels2twist(*a3, pu, P(tw))
        return f, A.reshape(3, 3).T, B.reshape(4, 3).T, tw

    @classmethod
    def normalize_angle_pi(cls, r):
        return cls.lib().eo_normalize_angle_pi(float(r))

    @classmethod
    def integrate_twist(cls, x, u, dt):
        x, px = _d(x); u, pu = _d(u); out = np.zeros(3)
        cls.lib().eo_integrate_twist(px, pu, dt, out.ctypes.data_as(_dp))
        return out

    @classmethod
    def rk4_forward(cls, model, dt, horizon, x0, ut):
        x0, p0 = _d(x0); ut, pu = _d(ut)
        steps = int(abs(horizon / dt)); xt = np.zeros((steps, 3))
        n = cls.lib().eo_rk4_forward(model, dt, horizon, p0, pu, xt.ctypes.data_as(_dp))
        if n < 0:
            raise ValueError("Invalid twist y-velocity must be 0.")
        return xt

    @classmethod
    def rk4_forward_cart(cls, r, b, dt, horizon, x0, ut):
        x0, p0 = _d(x0); ut, pu = _d(ut)
        steps = int(abs(horizon / dt)); xt = np.zeros((steps, 3))
        cls.lib().eo_rk4_forward_cart(r, b, dt, horizon, p0, pu, xt.ctypes.data_as(_dp))
        return xt

    @classmethod
    def rk4_forward_mecanum(cls, r, bx, by, dt, horizon, x0, ut):
        x0, p0 = _d(x0); ut, pu = _d(ut)
        steps = int(abs(horizon / dt)); xt = np.zeros((steps, 3))
        f = cls.lib().eo_rk4_forward_mecanum
        f.argtypes = [C.c_double] * 5 + [_dp, _dp, _dp]
        f(r, bx, by, dt, horizon, p0, pu, xt.ctypes.data_as(_dp))
        return xt

    @classmethod
    def entropy(cls, p):
        f = cls.lib().eo_entropy
        f.restype, f.argtypes = C.c_double, [C.c_double]
        return f(float(p))

    @classmethod
    def entropy_grid(cls, cells):
        """entropy(int8 / 100) of every cell (numerics.hpp:164-179 over grid.cpp:177-184)"""
        cells = np.ascontiguousarray(cells, dtype=np.int8)
        out = np.zeros(cells.shape)
        f = cls.lib().eo_entropy_grid
        f.argtypes = [C.c_void_p, C.c_longlong, _dp]
        f(cells.ctypes.data, cells.size, out.ctypes.data_as(_dp))
        return out

    @classmethod
    def phik_rows(cls, phi_rows, row_begin, res, lx, ly, nb, total):
        """un-normalised-by-rows partial of phik_from_grid: sum over the given rows of F_k * (phi / total)"""
        phi_rows, pp = _d(phi_rows); nrows, nx = phi_rows.shape
        acc = np.zeros(nb * nb)
        f = cls.lib().eo_phik_rows
        f.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, _dp]
        f(pp, nx, int(row_begin), nrows, res, lx, ly, nb, float(total), acc.ctypes.data_as(_dp))
        return acc

    @classmethod
    def basis_tables(cls, nb):
        k = np.zeros((nb * nb, 2), dtype=np.int64); lam = np.zeros(nb * nb)
        cls.lib().eo_basis_tables(C.c_int(nb), k.ctypes.data_as(C.POINTER(C.c_longlong)), lam.ctypes.data_as(_dp))
        return k, lam

    @classmethod
    def fourier_basis(cls, lx, ly, nb, x):
        x, px = _d(x); fk = np.zeros(nb * nb)
        cls.lib().eo_fourier_basis(C.c_double(lx), C.c_double(ly), C.c_int(nb), px, fk.ctypes.data_as(_dp))
        return fk

    @classmethod
    def grad_fourier_basis(cls, lx, ly, nb, x):
        x, px = _d(x); d = np.zeros((nb * nb, 2))
        cls.lib().eo_grad_fourier_basis(C.c_double(lx), C.c_double(ly), C.c_int(nb), px, d.ctypes.data_as(_dp))
        return d

    @classmethod
    def traj_coeff(cls, lx, ly, nb, xt):
        xt, p = _d(xt); ck = np.zeros(nb * nb)
        cls.lib().eo_traj_coeff(C.c_double(lx), C.c_double(ly), C.c_int(nb), p, C.c_int(xt.shape[1]),
                                C.c_int(xt.shape[0]), ck.ctypes.data_as(_dp))
        return ck

    @classmethod
    def target_grid(cls, lx, ly, res):
        nx, ny = C.c_int(0), C.c_int(0)
        cls.lib().eo_target_grid_dims(C.c_double(lx), C.c_double(ly), C.c_double(res), C.byref(nx), C.byref(ny))
        g = np.zeros((nx.value * ny.value, 2))
        cls.lib().eo_target_grid(C.c_double(res), nx, ny, g.ctypes.data_as(_dp))
        return g, nx.value, ny.value

    @classmethod
    def target_fill(cls, mu, sigma, trans, grid):
        mu, pm = _d(mu); sigma, ps = _d(sigma); trans, pt = _d(trans); grid, pg = _d(grid)
        vals = np.zeros(grid.shape[0])
        cls.lib().eo_target_fill(C.c_int(mu.size // 2), pm, ps, pt, pg, C.c_longlong(grid.shape[0]),
                                 vals.ctypes.data_as(_dp))
        return vals

    @classmethod
    def spatial_coeff(cls, lx, ly, nb, vals, grid):
        vals, pv = _d(vals); grid, pg = _d(grid); ph = np.zeros(nb * nb)
        cls.lib().eo_spatial_coeff(C.c_double(lx), C.c_double(ly), C.c_int(nb), pv, pg,
                                   C.c_longlong(vals.size), ph.ctypes.data_as(_dp))
        return ph

    @classmethod
    def phik_from_grid(cls, phi, res, lx, ly, nb):
        phi, pp = _d(phi); ny, nx = phi.shape
        ph = np.zeros(nb * nb); s = C.c_double(0.0)
        cls.lib().eo_phik_from_grid(pp, nx, ny, res, lx, ly, nb, ph.ctypes.data_as(_dp),
                                    C.cast(C.byref(s), _dp))
        return ph, s.value


    # ---- occupancy-grid collision checking ------------
    class _Grid(C.Structure):
        _fields_ = [("data", C.c_void_p), ("xsize", C.c_uint), ("ysize", C.c_uint),
                    ("resolution", C.c_double), ("xmin", C.c_double), ("ymin", C.c_double)]

    class _Collision(C.Structure):
        _fields_ = [("boundary_radius", C.c_double), ("search_radius", C.c_double),
                    ("obstacle_threshold", C.c_double), ("occupied_threshold", C.c_double)]

    @classmethod
    def _grid_args(cls, data, res, xmin, ymin, col):
        data = np.ascontiguousarray(data, dtype=np.int8)
        g = cls._Grid(data.ctypes.data, data.shape[1], data.shape[0], res, xmin, ymin)
        c = cls._Collision(*[float(v) for v in col])
        return data, g, c

    @classmethod
    def collision_check(cls, data, res, xmin, ymin, col, poses):
        """col = (boundary_radius, search_radius, obstacle_threshold, occupied_threshold); -> int array, 1 = hit"""
        data, g, c = cls._grid_args(data, res, xmin, ymin, col)
        poses, _ = _d(poses)
        poses = poses.reshape(-1, 3)
        f = cls.lib().eo_collision_check
        f.argtypes = [C.c_void_p, C.c_void_p, _dp]
        return np.array([f(C.byref(g), C.byref(c), poses[i].ctypes.data_as(_dp)) for i in range(len(poses))],
                        dtype=np.int32)

    @classmethod
    def validate_control(cls, data, res, xmin, ymin, col, x0, u, dt, horizon):
        """-> int array, 1 = collision free (numerics.hpp:312-330)"""
        data, g, c = cls._grid_args(data, res, xmin, ymin, col)
        x0, _ = _d(x0); u, _ = _d(u)
        x0, u = x0.reshape(-1, 3), u.reshape(-1, 3)
        f = cls.lib().eo_validate_control
        f.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, C.c_double, C.c_double]
        return np.array([f(C.byref(g), C.byref(c), x0[i].ctypes.data_as(_dp), u[i].ctypes.data_as(_dp), dt, horizon)
                         for i in range(len(x0))], dtype=np.int32)


    # ---- DynamicWindow --------------------------------
    class _Dwa(C.Structure):
        _fields_ = [(n, C.c_double) for n in ("dt", "horizon", "acc_dt", "acc_lim_x", "acc_lim_y", "acc_lim_th",
                                              "max_vel_x", "min_vel_x", "max_vel_y", "min_vel_y", "max_rot_vel",
                                              "min_rot_vel")] + [(n, C.c_uint) for n in ("vx", "vy", "vth")]

    @classmethod
    def dwa_control(cls, data, res, xmin, ymin, col, dwa_cfg, samples, x0, vb, vref=None, xt_ref=None, dt_ref=0.1):
        """dwa_cfg = the 12 doubles of the DynamicWindow constructor after `collision`
        (dynamic_window.hpp:60-65), samples = (vx, vy, vth).  vref: (B, 3) reference twists, or
        xt_ref: (ncols, 3) one reference trajectory.  -> found (B,), u_opt (B, 3), min_cost (B,)"""
        data, g, c = cls._grid_args(data, res, xmin, ymin, col)
        d = cls._Dwa(*[float(v) for v in dwa_cfg], *[int(v) for v in samples])
        x0, _ = _d(x0); vb, _ = _d(vb)
        x0, vb = x0.reshape(-1, 3), vb.reshape(-1, 3)
        n = len(x0)
        found, u, cost = np.zeros(n, dtype=np.int32), np.zeros((n, 3)), np.zeros(n)
        lib = cls.lib()
        if vref is not None:
            vref, _ = _d(vref); vref = vref.reshape(-1, 3)
            f = lib.eo_dwa_control_twist
            f.argtypes = [C.c_void_p] * 3 + [_dp] * 5
            for i in range(n):
                found[i] = f(C.byref(g), C.byref(c), C.byref(d), x0[i].ctypes.data_as(_dp), vb[i].ctypes.data_as(_dp),
                             vref[i].ctypes.data_as(_dp), u[i].ctypes.data_as(_dp), cost[i:].ctypes.data_as(_dp))
        else:
            xt, pxt = _d(xt_ref)
            f = lib.eo_dwa_control_traj
            f.argtypes = [C.c_void_p] * 3 + [_dp] * 3 + [C.c_int, C.c_double, _dp, _dp]
            for i in range(n):
                found[i] = f(C.byref(g), C.byref(c), C.byref(d), x0[i].ctypes.data_as(_dp), vb[i].ctypes.data_as(_dp),
                             pxt, xt.shape[0], float(dt_ref), u[i].ctypes.data_as(_dp), cost[i:].ctypes.data_as(_dp))
        return found, u, cost


class RefLib:
    """The unmodified reference, compiled against the shim."""

    _lib = None

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    @classmethod
    def lib(cls):
        if cls._lib is None:
            lib = C.CDLL(REF_SO)
            lib.ref_last_error.restype = C.c_char_p
            lib.ref_normalize_angle_pi.restype = C.c_double
            lib.ref_normalize_angle_pi.argtypes = [C.c_double]
            lib.ref_create.restype = C.c_void_p
            lib.ref_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                       C.c_longlong, C.c_int, _dp, _dp, _dp]
            lib.ref_destroy.argtypes = [C.c_void_p]
            lib.ref_set_target.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
            lib.ref_add_state_memory.argtypes = [C.c_void_p, _dp]
            lib.ref_memory_size.argtypes = [C.c_void_p]
            lib.ref_memory_size.restype = C.c_longlong
            lib.ref_steps.argtypes = [C.c_void_p]
            ctl = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp]
            lib.ref_control.argtypes = ctl
            lib.ref_control_trace.argtypes = ctl
            lib.ref_get_last.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp]
            lib.ref_config_target.argtypes = [C.c_void_p] + [C.c_double] * 5
            lib.ref_opt_traj.argtypes = [C.c_void_p, _dp]
            lib.ref_get_ut.argtypes = [C.c_void_p, _dp]
            lib.ref_set_ut.argtypes = [C.c_void_p, _dp]
            lib.ref_get_phik.argtypes = [C.c_void_p, _dp, _dp]
            lib.ref_control_many.argtypes = [C.POINTER(C.c_void_p), C.c_int] + [C.c_double] * 5 + [_dp, _dp]
            lib.ref_rk4_forward.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp, _dp]
            lib.ref_rk4_forward_cart.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp]
            lib.ref_integrate_twist.argtypes = [_dp, _dp, C.c_double, _dp]
            cls._lib = lib
        return cls._lib

    @classmethod
    def create(cls, model, dt, horizon, resolution, expl_weight, num_basis, buffer_size, batch_size,
               Rinv, umin, umax, map_res=0.1):
        lib = cls.lib()
        Rinv, pR = _d(np.asarray(Rinv, dtype=np.float64).T)
        umin, pmin = _d(umin)
        umax, pmax = _d(umax)
        h = lib.ref_create(model, dt, horizon, resolution, expl_weight, num_basis, buffer_size, batch_size,
                           pR, pmin, pmax)
        if not h:
            raise ValueError(lib.ref_last_error().decode())
        steps = lib.ref_steps(C.c_void_p(h))
        return _Controller(lib, "ref_", C.c_void_p(h), num_basis, steps, True, map_res)

    @classmethod
    def model_f(cls, model, x, u):
        x, px = _d(x); u, pu = _d(u); out = np.zeros(3)
        if cls.lib().ref_model_f(C.c_int(model), px, pu, out.ctypes.data_as(_dp)):
            raise ValueError(cls.lib().ref_last_error().decode())
        return out

    @classmethod
    def model_fdx(cls, model, x, u):
        x, px = _d(x); u, pu = _d(u); A = np.zeros(9)
        cls.lib().ref_model_fdx(C.c_int(model), px, pu, A.ctypes.data_as(_dp))
        return A.reshape(3, 3).T

    @classmethod
    def model_fdu(cls, model, x):
        x, px = _d(x); B = np.zeros(9)
        cls.lib().ref_model_fdu(C.c_int(model), px, B.ctypes.data_as(_dp))
        return B.reshape(3, 3).T

    @classmethod
    def cart(cls, r, b, x, u):
        x, px = _d(x); u, pu = _d(u)
        f, A, B, tw = np.zeros(3), np.zeros(9), np.zeros(6), np.zeros(3)
        P = lambda a: a.ctypes.data_as(_dp)
        cls.lib().ref_cart(C.c_double(r), C.c_double(b), px, pu, P(f), P(A), P(B), P(tw))
        return f, A.reshape(3, 3).T, B.reshape(2, 3).T, tw

    @classmethod
    def mecanum(cls, r, bx, by, x, u):
        x, px = _d(x); u, pu = _d(u)
        f, A, B, tw = np.zeros(3), np.zeros(9), np.zeros(12), np.zeros(3)
        P = lambda a: a.ctypes.data_as(_dp)
        cls.lib().ref_mecanum(C.c_double(r), C.c_double(bx), C.c_double(by), px, pu, P(f), P(A), P(B), P(tw))
        return f, A.reshape(3, 3).T, B.reshape(4, 3).T, tw

    @classmethod
    def normalize_angle_pi(cls, r):
        return cls.lib().ref_normalize_angle_pi(float(r))

    @classmethod
    def integrate_twist(cls, x, u, dt):
        x, px = _d(x); u, pu = _d(u); out = np.zeros(3)
        cls.lib().ref_integrate_twist(px, pu, dt, out.ctypes.data_as(_dp))
        return out

    @classmethod
    def rk4_forward(cls, model, dt, horizon, x0, ut):
        x0, p0 = _d(x0); ut, pu = _d(ut)
        steps = int(abs(horizon / dt)); xt = np.zeros((steps, 3))
        if cls.lib().ref_rk4_forward(model, dt, horizon, p0, pu, xt.ctypes.data_as(_dp)) < 0:
            raise ValueError(cls.lib().ref_last_error().decode())
        return xt

    @classmethod
    def rk4_forward_cart(cls, r, b, dt, horizon, x0, ut):
        x0, p0 = _d(x0); ut, pu = _d(ut)
        steps = int(abs(horizon / dt)); xt = np.zeros((steps, 3))
        cls.lib().ref_rk4_forward_cart(r, b, dt, horizon, p0, pu, xt.ctypes.data_as(_dp))
        return xt

    @classmethod
    def rk4_forward_mecanum(cls, r, bx, by, dt, horizon, x0, ut):
        x0, p0 = _d(x0); ut, pu = _d(ut)
        steps = int(abs(horizon / dt)); xt = np.zeros((steps, 3))
        f = cls.lib().ref_rk4_forward_mecanum
        f.argtypes = [C.c_double] * 5 + [_dp, _dp, _dp]
        f(r, bx, by, dt, horizon, p0, pu, xt.ctypes.data_as(_dp))
        return xt

    @classmethod
    def entropy(cls, p):
        f = cls.lib().ref_entropy
        f.restype, f.argtypes = C.c_double, [C.c_double]
        return f(float(p))

    @classmethod
    def entropy_grid(cls, cells, res=0.05):
        """entropy(GridMap::getCell(i)) for every cell, through the compiled reference"""
        cells = np.ascontiguousarray(cells, dtype=np.int8)
        ysize, xsize = cells.shape
        out = np.zeros(cells.shape)
        f = cls.lib().ref_entropy_grid
        f.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_double, _dp]
        f(cells.ctypes.data, xsize, ysize, float(res), out.ctypes.data_as(_dp))
        return out

    @classmethod
    def basis_tables(cls, nb):
        k = np.zeros((nb * nb, 2), dtype=np.int64); lam = np.zeros(nb * nb)
        cls.lib().ref_basis_tables(C.c_int(nb), k.ctypes.data_as(C.POINTER(C.c_longlong)), lam.ctypes.data_as(_dp))
        return k, lam

    @classmethod
    def fourier_basis(cls, lx, ly, nb, x):
        x, px = _d(x); fk = np.zeros(nb * nb)
        cls.lib().ref_fourier_basis(C.c_double(lx), C.c_double(ly), C.c_int(nb), px, fk.ctypes.data_as(_dp))
        return fk

    @classmethod
    def grad_fourier_basis(cls, lx, ly, nb, x):
        x, px = _d(x); d = np.zeros((nb * nb, 2))
        cls.lib().ref_grad_fourier_basis(C.c_double(lx), C.c_double(ly), C.c_int(nb), px, d.ctypes.data_as(_dp))
        return d

    @classmethod
    def traj_coeff(cls, lx, ly, nb, xt):
        xt, p = _d(xt); ck = np.zeros(nb * nb)
        cls.lib().ref_traj_coeff(C.c_double(lx), C.c_double(ly), C.c_int(nb), p, C.c_int(xt.shape[1]),
                                 C.c_int(xt.shape[0]), ck.ctypes.data_as(_dp))
        return ck

    @classmethod
    def target_fill(cls, mu, sigma, trans, grid):
        mu, pm = _d(mu); sigma, ps = _d(sigma); trans, pt = _d(trans); grid, pg = _d(grid)
        vals = np.zeros(grid.shape[0])
        cls.lib().ref_target_fill(C.c_int(mu.size // 2), pm, ps, pt, pg, C.c_longlong(grid.shape[0]),
                                  vals.ctypes.data_as(_dp))
        return vals

    @classmethod
    def spatial_coeff(cls, lx, ly, nb, vals, grid):
        vals, pv = _d(vals); grid, pg = _d(grid); ph = np.zeros(nb * nb)
        cls.lib().ref_spatial_coeff(C.c_double(lx), C.c_double(ly), C.c_int(nb), pv, pg,
                                    C.c_longlong(vals.size), ph.ctypes.data_as(_dp))
        return ph

    # ---- occupancy-grid collision checking: the reference's own GridMap / Collision ----
    @classmethod
    def collision_check(cls, data, res, xmin, ymin, col, poses):
        data = np.ascontiguousarray(data, dtype=np.int8)
        poses, pp = _d(poses)
        n = poses.size // 3
        hit = np.zeros(n, dtype=np.int32)
        f = cls.lib().ref_collision_check_many
        f.argtypes = [C.c_void_p, C.c_uint, C.c_uint] + [C.c_double] * 7 + [_dp, C.c_int, _ip]
        if f(data.ctypes.data, data.shape[1], data.shape[0], res, xmin, ymin, *[float(v) for v in col], pp, n,
             hit.ctypes.data_as(_ip)) != 0:
            raise ValueError(cls.lib().ref_last_error().decode())
        return hit

    @classmethod
    def validate_control(cls, data, res, xmin, ymin, col, x0, u, dt, horizon):
        data = np.ascontiguousarray(data, dtype=np.int8)
        x0, px = _d(x0); u, pu = _d(u)
        n = x0.size // 3
        valid = np.zeros(n, dtype=np.int32)
        f = cls.lib().ref_validate_control_many
        f.argtypes = [C.c_void_p, C.c_uint, C.c_uint] + [C.c_double] * 7 + [_dp, _dp, C.c_int, C.c_double,
                                                                             C.c_double, _ip]
        if f(data.ctypes.data, data.shape[1], data.shape[0], res, xmin, ymin, *[float(v) for v in col], px, pu, n,
             dt, horizon, valid.ctypes.data_as(_ip)) != 0:
            raise ValueError(cls.lib().ref_last_error().decode())
        return valid

    @classmethod
    def dwa_control(cls, data, res, xmin, ymin, col, dwa_cfg, samples, x0, vb, vref=None, xt_ref=None, dt_ref=0.1):
        """the reference's own DynamicWindow::control (both overloads); -> found (B,), u_opt (B, 3)"""
        data = np.ascontiguousarray(data, dtype=np.int8)
        x0, px = _d(x0); vb, pv = _d(vb)
        n = x0.size // 3
        colv, pc = _d(col); cfg, pcfg = _d(dwa_cfg)
        smp = np.ascontiguousarray(samples, dtype=np.uint32)
        found, u = np.zeros(n, dtype=np.int32), np.zeros((n, 3))
        f = cls.lib().ref_dwa_control_many
        f.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, _dp, _dp, C.c_void_p, _dp, _dp,
                      C.c_int, C.c_int, _dp, C.c_int, C.c_double, _ip, _dp]
        if vref is not None:
            ref, pr = _d(vref); mode, ncols = 0, 0
        else:
            ref, pr = _d(xt_ref); mode, ncols = 1, ref.shape[0]
        if f(data.ctypes.data, data.shape[1], data.shape[0], res, xmin, ymin, pc, pcfg, smp.ctypes.data, px, pv, n, mode,
             pr, ncols, float(dt_ref), found.ctypes.data_as(_ip), u.ctypes.data_as(_dp)) != 0:
            raise ValueError(cls.lib().ref_last_error().decode())
        return found, u
