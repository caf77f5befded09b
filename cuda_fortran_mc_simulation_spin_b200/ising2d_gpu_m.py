"""Host mirror of ``module ising2d_gpu_m`` (src/ising2d_gpu_m.f90): same type
and procedure names as ``type(ising2d_gpu)`` (:12-42), over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._ising_base import HEATBATH, METROPOLIS, _IsingBase  # noqa: F401
from ._lib import P, PP, f64, i32, i64

ising2d_gpu_stat = 0  # mirrors `integer(int32), public, protected :: ising2d_gpu_stat` (:8)


class ising2d_gpu(_IsingBase):
    _pfx = "b200mc_ising2d"
    _ndim = 2

    def init(self, nx, ny, kbt, iseed):
        """init_ising2d_gpu(this, nx, ny, kbt, iseed), :44-61"""
        if self._h:
            self._f("destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        f = _lib.fn("b200mc_ising2d_create", C.c_int, PP, i64, i64, f64, i32)
        _lib.check(f(C.byref(self._h), int(nx), int(ny), float(kbt), int(iseed)))
        return self

    def init_packed(self, nx, ny, kbt, iseed):
        """init on the bit-packed (multi-spin coded) storage: one bit per site, Metropolis, one GPU"""
        return self._init_packed((nx, ny), kbt, iseed)

    def init_packed_distributed(self, nx, ny, kbt, iseed, group=None):
        """the global lattice on bit-packed storage, one slab per rank of the torch.distributed group"""
        return self._init_packed_distributed((nx, ny), kbt, iseed, group)

    def init_slab(self, nx, ny, kbt, iseed, rank, nranks, nccl_id):
        """the global nx x ny lattice, this process owning slab `rank` of `nranks` (one GPU each)"""
        return self._init_slab((nx, ny), kbt, iseed, rank, nranks, nccl_id)

    def init_distributed(self, nx, ny, kbt, iseed, group=None):
        return self._init_torch_distributed((nx, ny), kbt, iseed, group)

    def _halo(self):
        return self.nx()

    def exparr(self):
        """host copy of exparr(-8:8) (:126-130): exparr()[d + 8]"""
        out = np.empty(17, dtype=np.float64)
        self._call("get_exparr", out.ctypes.data_as(P), argtypes=(P,))
        return out
