"""Host mirror of ``module clock_gpu_m`` (src/clock_gpu_m.f90): ``type(clock_gpu)`` (:13-47)
with the reference's procedure names, over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import P, PP, f64, i32, i64

clock_gpu_stat = 0  # mirrors `integer(int32), public, protected :: clock_gpu_stat` (:8)
clock_max_state_limit = 16  # reference: 50 (:10); the q^6 class table bounds it here


class clock_gpu:
    _multi = False

    def __init__(self):
        self._h = C.c_void_p(None)

    def _call(self, name, *args, argtypes=()):
        f = _lib.fn(f"b200mc_clock_{name}", C.c_int, P, *argtypes)
        _lib.check(f(self._h, *args))

    def _get(self, name, restype):
        return _lib.fn(f"b200mc_clock_{name}", restype, P)(self._h)

    def __del__(self):
        try:
            if self._h:
                _lib.fn("b200mc_clock_destroy", C.c_int, P)(self._h)
                self._h = C.c_void_p(None)
        except Exception:
            pass

    def init(self, nx, ny, kbt, state, iseed):
        """init_clock_gpu(this, nx, ny, kbt, state, iseed), :49-79"""
        if self._h:
            _lib.fn("b200mc_clock_destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        f = _lib.fn("b200mc_clock_create", C.c_int, PP, i64, i64, f64, i32, i32)
        _lib.check(f(C.byref(self._h), int(nx), int(ny), float(kbt), int(state), int(iseed)))
        return self

    def set_allup_spin(self): self._call("set_allup_spin")
    def set_random_spin(self): self._call("set_random_spin")
    def set_kbt(self, kbt): self._call("set_kbt", float(kbt), argtypes=(f64,))
    def set_beta(self, beta): self._call("set_beta", float(beta), argtypes=(f64,))
    def skip_curand(self, n_skip): self._call("skip_curand", int(n_skip), argtypes=(i64,))
    def update(self): self._call("update")
    def update_n(self, n): self._call("update_n", int(n), argtypes=(i32,))

    def update_with_randoms(self, randoms, next_states):
        r = np.ascontiguousarray(randoms, dtype=np.float64)
        p = np.ascontiguousarray(next_states, dtype=np.float64)
        n = self.nall() * self.n_multi()
        if r.size != n or p.size != n:
            raise ValueError("randoms / next_states must hold nall (x n_multi) uniforms each")
        self._call("update_with_randoms", r.ctypes.data_as(P), p.ctypes.data_as(P), argtypes=(P, P))

    def nx(self): return int(self._get("nx", i64))
    def ny(self): return int(self._get("ny", i64))
    def nall(self): return int(self._get("nall", i64))
    def state(self): return int(self._get("state", i32))
    def n_multi(self): return int(self._get("n_multi", i32))
    def kbt(self): return float(self._get("kbt", f64))
    def beta(self): return float(self._get("beta", f64))
    def sync(self): self._call("sync")

    def spins(self):
        out = np.empty((self.n_multi(), self.nall() + 2 * self.nx()), dtype=np.int32)
        self._call("get_spins", out.ctypes.data_as(P), argtypes=(P,))
        return out if self._multi else out[0]

    def set_spins(self, spins):
        s = np.ascontiguousarray(spins, dtype=np.int32)
        if s.size != self.n_multi() * (self.nall() + 2 * self.nx()):
            raise ValueError("spins must use the reference layout, halo cells included")
        self._call("set_spins", s.ctypes.data_as(P), argtypes=(P,))

    def ws(self):
        q = self.state()
        out = np.empty(q ** 6, dtype=np.float64)
        self._call("get_ws", out.ctypes.data_as(P), argtypes=(P,))
        return out

    def histograms(self):
        """exact integer observables: (hist, bond_left, bond_down), each (n_multi, q)"""
        q, n = self.state(), self.n_multi()
        h = np.zeros((n, q), dtype=np.int64)
        bl = np.zeros((n, q), dtype=np.int64)
        bd = np.zeros((n, q), dtype=np.int64)
        self._call("get_histograms", h.ctypes.data_as(P), bl.ctypes.data_as(P), bd.ctypes.data_as(P), argtypes=(P, P, P))
        return h, bl, bd

    def _obs(self, name):
        res = np.zeros(self.n_multi(), dtype=np.float64)
        self._call(name, res.ctypes.data_as(P), argtypes=(P,))
        return res

    def calc_energy_sum(self):
        """:245-262 (real64)"""
        return float(self._obs("calc_energy_sum")[0])

    def calc_magne_sum(self):
        """:264-280 (real64)"""
        return float(self._obs("calc_magne_sum")[0])
