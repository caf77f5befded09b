"""Shared host-side plumbing of the two helical Ising mirrors (ctypes over the C ABI)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import P, PP, f64, i32, i64

METROPOLIS = 0
HEATBATH = 1


class _IsingBase:
    _pfx = ""      # "b200mc_ising2d" / "b200mc_ising3d"
    _ndim = 0

    def __init__(self):
        self._h = C.c_void_p(None)

    # -- plumbing ---------------------------------------------------------
    def _f(self, name, restype, *argtypes):
        return _lib.fn(f"{self._pfx}_{name}", restype, *argtypes)

    def _call(self, name, *args, argtypes=()):
        f = self._f(name, C.c_int, P, *argtypes)
        _lib.check(f(self._h, *args))

    def __del__(self):
        try:
            if self._h:
                self._f("destroy", C.c_int, P)(self._h)
                self._h = C.c_void_p(None)
        except Exception:
            pass

    # -- setters (reference names) ---------------------------------------
    def set_allup_spin(self):
        self._call("set_allup_spin")

    def set_random_spin(self):
        self._call("set_random_spin")

    def set_kbt(self, kbt):
        self._call("set_kbt", float(kbt), argtypes=(f64,))

    def set_beta(self, beta):
        self._call("set_beta", float(beta), argtypes=(f64,))

    def skip_curand(self, n_skip):
        self._call("skip_curand", int(n_skip), argtypes=(i64,))

    def set_method(self, method):
        self._call("set_method", int(method), argtypes=(i32,))

    # -- updaters -----------------------------------------------------------
    def update(self):
        self._call("update")

    def update_n(self, n_sweeps):
        self._call("update_n", int(n_sweeps), argtypes=(i32,))

    def update_with_randoms(self, randoms):
        r = np.ascontiguousarray(randoms, dtype=np.float64)
        if r.size != self.nall():
            raise ValueError(f"randoms must hold nall = {self.nall()} uniforms")
        self._call("update_with_randoms", r.ctypes.data_as(P), argtypes=(P,))

    # -- getters ------------------------------------------------------------
    def nx(self):
        return int(self._f("nx", i64, P)(self._h))

    def ny(self):
        return int(self._f("ny", i64, P)(self._h))

    def nall(self):
        return int(self._f("nall", i64, P)(self._h))

    def kbt(self):
        return float(self._f("kbt", f64, P)(self._h))

    def beta(self):
        return float(self._f("beta", f64, P)(self._h))

    def _halo(self):
        raise NotImplementedError

    def spins(self):
        out = np.empty(self.nall() + 2 * self._halo(), dtype=np.int32)
        self._call("get_spins", out.ctypes.data_as(P), argtypes=(P,))
        return out

    def set_spins(self, spins):
        s = np.ascontiguousarray(spins, dtype=np.int32)
        if s.size != self.nall() + 2 * self._halo():
            raise ValueError("spins must use the reference layout, halo cells included")
        self._call("set_spins", s.ctypes.data_as(P), argtypes=(P,))

    # -- calculators ----------------------------------------------------------
    def calc_energy_sum(self):
        e = C.c_int64(0)
        self._call("calc_energy_sum", C.byref(e), argtypes=(C.POINTER(C.c_int64),))
        return int(e.value)

    def calc_magne_sum(self):
        m = C.c_int64(0)
        self._call("calc_magne_sum", C.byref(m), argtypes=(C.POINTER(C.c_int64),))
        return int(m.value)

    def measure(self):
        e, m = C.c_int64(0), C.c_int64(0)
        self._call("measure", C.byref(e), C.byref(m), argtypes=(C.POINTER(C.c_int64), C.POINTER(C.c_int64)))
        return int(e.value), int(m.value)

    def sync(self):
        self._call("sync")

    def set_timing(self, on=True):
        """per-launch CUDA-event timing of the colour-pass kernel (bench.py roofline leg)"""
        self._call("set_timing", 1 if on else 0, argtypes=(i32,))

    def get_timing(self):
        n, ms = C.c_int64(0), C.c_double(0.0)
        self._call("get_timing", C.byref(n), C.byref(ms), argtypes=(C.POINTER(C.c_int64), C.POINTER(C.c_double)))
        return int(n.value), float(ms.value)

    def set_stream(self, cuda_stream: int):
        self._call("set_stream", C.c_void_p(cuda_stream), argtypes=(P,))
