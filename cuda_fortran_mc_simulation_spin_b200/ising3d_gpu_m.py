"""Host mirror of ``module ising3d_gpu_m`` (src/ising3d_gpu_m.f90): same type
and procedure names as ``type(ising3d_gpu)`` (:15-48), over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._ising_base import HEATBATH, METROPOLIS, _IsingBase  # noqa: F401
from ._lib import P, PP, f64, i32, i64

ising3d_gpu_stat = 0  # mirrors `integer(int32), public, protected :: ising3d_gpu_stat` (:8)


class ising3d_gpu(_IsingBase):
    _pfx = "b200mc_ising3d"
    _ndim = 3

    def init(self, nx, ny, nz, kbt, iseed):
        """init_ising3d_gpu(this, nx, ny, nz, kbt, iseed), :50-71"""
        if self._h:
            self._f("destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        f = _lib.fn("b200mc_ising3d_create", C.c_int, PP, i64, i64, i64, f64, i32)
        _lib.check(f(C.byref(self._h), int(nx), int(ny), int(nz), float(kbt), int(iseed)))
        return self

    def init_packed(self, nx, ny, nz, kbt, iseed):
        """init on the bit-packed (multi-spin coded) storage: one bit per site, Metropolis, one GPU"""
        return self._init_packed((nx, ny, nz), kbt, iseed)

    def init_packed_distributed(self, nx, ny, nz, kbt, iseed, group=None):
        """the global lattice on bit-packed storage, one slab per rank of the torch.distributed group"""
        return self._init_packed_distributed((nx, ny, nz), kbt, iseed, group)

    def init_slab(self, nx, ny, nz, kbt, iseed, rank, nranks, nccl_id):
        """the global nx x ny x nz lattice, this process owning slab `rank` of `nranks` (one GPU each)"""
        return self._init_slab((nx, ny, nz), kbt, iseed, rank, nranks, nccl_id)

    def init_distributed(self, nx, ny, nz, kbt, iseed, group=None):
        return self._init_torch_distributed((nx, ny, nz), kbt, iseed, group)

    def nz(self):
        return int(self._f("nz", i64, P)(self._h))

    def _halo(self):
        return self.nx() * self.ny()

    def ws(self):
        """host copy of ws(0:6, 0:1) (:153-171), shape (2, 7): ws()[s, S]"""
        out = np.empty(14, dtype=np.float64)
        self._call("get_ws", out.ctypes.data_as(P), argtypes=(P,))
        return out.reshape(2, 7)
