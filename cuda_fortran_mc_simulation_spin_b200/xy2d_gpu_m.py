"""Host mirror of ``module xy2d_gpu_m`` (src/xy2d_gpu_m.f90): ``type(xy2d_gpu)`` (:12-43), the XY model with the
helical boundary / linear-index colouring of ``ising2d_gpu_m``, over the C ABI (``b200mc_xy2dh_*``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import P, PP, f64, i32, i64

xy2d_gpu_stat = 0  # mirrors `integer(int32), public, protected :: xy2d_gpu_stat` (:8)


class xy2d_gpu:
    def __init__(self):
        self._h = C.c_void_p(None)

    def _call(self, name, *args, argtypes=()):
        f = _lib.fn(f"b200mc_xy2dh_{name}", C.c_int, P, *argtypes)
        _lib.check(f(self._h, *args))

    def _get(self, name, restype):
        return _lib.fn(f"b200mc_xy2dh_{name}", restype, P)(self._h)

    def _dbl(self, name):
        r = C.c_double(0.0)
        self._call(name, C.byref(r), argtypes=(C.POINTER(C.c_double),))
        return float(r.value)

    def __del__(self):
        try:
            if self._h:
                _lib.fn("b200mc_xy2dh_destroy", C.c_int, P)(self._h)
                self._h = C.c_void_p(None)
        except Exception:
            pass

    def init(self, nx, ny, kbt, iseed):
        """init_xy2d_gpu(this, nx, ny, kbt, iseed), :45-62"""
        if self._h:
            _lib.fn("b200mc_xy2dh_destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        f = _lib.fn("b200mc_xy2dh_create", C.c_int, PP, i64, i64, f64, i32)
        _lib.check(f(C.byref(self._h), int(nx), int(ny), float(kbt), int(iseed)))
        return self

    def skip_curand(self, n_skip): self._call("skip_curand", int(n_skip), argtypes=(i64,))
    def set_allup_spin(self): self._call("set_allup_spin")
    def set_random_spin(self): self._call("set_random_spin")
    def set_kbt(self, kbt): self._call("set_kbt", float(kbt), argtypes=(f64,))
    def set_beta(self, beta): self._call("set_beta", float(beta), argtypes=(f64,))
    def update(self): self._call("update")
    def update_n(self, n): self._call("update_n", int(n), argtypes=(i32,))
    def update_over_relaxation(self, n_steps): self._call("update_over_relaxation", int(n_steps), argtypes=(i32,))
    def nx(self): return int(self._get("nx", i64))
    def ny(self): return int(self._get("ny", i64))
    def nall(self): return int(self._get("nall", i64))
    def kbt(self): return float(self._get("kbt", f64))
    def beta(self): return float(self._get("beta", f64))
    def sync(self): self._call("sync")
    def calc_energy_sum(self): return self._dbl("calc_energy_sum")
    def calc_magne_sum(self): return self._dbl("calc_magne_sum")

    def spins(self):
        """real64 (cos, sin) in the reference layout spins(1-nx : nall+nx, 1:2): array [2][nall + 2 nx]"""
        out = np.empty((2, self.nall() + 2 * self.nx()), dtype=np.float64)
        self._call("get_spins", out.ctypes.data_as(P), argtypes=(P,))
        return out

    def angles(self):
        """native state: fp32 angles in turns, linear index order"""
        out = np.empty(self.nall(), dtype=np.float32)
        self._call("get_angles", out.ctypes.data_as(P), argtypes=(P,))
        return out

    def set_angles(self, turns):
        t = np.ascontiguousarray(turns, dtype=np.float32)
        if t.size != self.nall():
            raise ValueError("angles must hold nall values")
        self._call("set_angles", t.ctypes.data_as(P), argtypes=(P,))
