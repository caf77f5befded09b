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

    def set_sample_offset(self, first_sample):
        """this handle's replicas are replicas first_sample .. first_sample + n_multi - 1 of the job (a batch split across
        GPUs, one handle per rank; call right after init)"""
        _lib.check(_lib.fn("b200mc_clock_set_sample_offset", C.c_int, C.c_void_p, i32)(self._h, int(first_sample)))
        return self

    def init_distributed(self, nx, ny, kbt, state, n_multi, iseed, group=None):
        """The n_multi replicas of the job shared out over the ranks of a torch.distributed group -- the replicas are
        independent (src/clock_gpu_multi_m.f90:215-236), so there is no exchange on the data path; calc_*_sum_all gather
        the per-replica observables of the whole job.  Every replica draws the stream it has in a one-GPU batch."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        lo, hi = split_samples(n_multi, rank, world)
        if hi == lo:
            raise ValueError(f"rank {rank}: no replica to run ({n_multi} replicas on {world} ranks)")
        self.init(nx, ny, kbt, state, hi - lo, iseed)
        self.set_sample_offset(lo)
        self._dist = (dist, group, n_multi)
        return self

    def _gather(self, local):
        import torch
        dist, group, total = self._dist
        world = dist.get_world_size(group)
        parts = [None] * world
        dist.all_gather_object(parts, [float(x) for x in local], group=group)
        return np.array([x for p in parts for x in p], dtype=np.float64)

    def calc_energy_sum_all(self):
        return self._gather(self.calc_energy_sum())

    def calc_magne_sum_all(self):
        return self._gather(self.calc_magne_sum())

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


def split_samples(n_multi, rank, world):
    """contiguous share [lo, hi) of n_multi samples for `rank` of `world` (the first n_multi % world ranks get one more)"""
    base, rem = divmod(int(n_multi), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
