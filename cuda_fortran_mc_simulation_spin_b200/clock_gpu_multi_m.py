"""Host mirror of ``module clock_gpu_multi_m`` (src/clock_gpu_multi_m.f90): the batched twin of
``clock_gpu_m`` -- ``init(nx, ny, kbt, state, n_multi, iseed)`` (:50-83), strict accept test
(:230-235), and ``calc_*_sum(res)`` filling a per-replica array (:265-315)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PP, f64, i32, i64
from .clock_gpu_m import clock_gpu as _single


class clock_gpu(_single):
    _multi = True

    def init(self, nx, ny, kbt, state, n_multi, iseed):
        if self._h:
            _lib.fn("b200mc_clock_destroy", C.c_int, C.c_void_p)(self._h)
            self._h = C.c_void_p(None)
        f = _lib.fn("b200mc_clock_multi_create", C.c_int, PP, i64, i64, f64, i32, i32, i32)
        _lib.check(f(C.byref(self._h), int(nx), int(ny), float(kbt), int(state), int(n_multi), int(iseed)))
        return self

    def calc_energy_sum(self, res=None):
        out = self._obs("calc_energy_sum")
        if res is not None:
            res[:] = out
        return out

    def calc_magne_sum(self, res=None):
        out = self._obs("calc_magne_sum")
        if res is not None:
            res[:] = out
        return out
