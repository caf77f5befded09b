"""Handle class shared by the two periodic-clock mirrors (clock_tableall_gpu_m,
clock_dual_lattice_tableall_gpu_m): one C-ABI object, ``b200mc_sixclock_*``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import P, PP, f64, i32, i64


class sixclock:
    def __init__(self, nx, ny, kbt, mstate=6, n_multi=1, iseed=42, variant=0):
        self._h = C.c_void_p(None)
        f = _lib.fn("b200mc_sixclock_create_variant", C.c_int, PP, i64, i64, f64, i32, i32, i32, i32)
        _lib.check(f(C.byref(self._h), int(nx), int(ny), float(kbt), int(mstate), int(n_multi), int(iseed), int(variant)))

    def _call(self, name, *args, argtypes=()):
        f = _lib.fn(f"b200mc_sixclock_{name}", C.c_int, P, *argtypes)
        _lib.check(f(self._h, *args))

    def _get(self, name, restype):
        return _lib.fn(f"b200mc_sixclock_{name}", restype, P)(self._h)

    def close(self):
        if self._h:
            _lib.fn("b200mc_sixclock_destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def nx(self): return int(self._get("nx", i64))
    def ny(self): return int(self._get("ny", i64))
    def nall(self): return int(self._get("nall", i64))
    def mstate(self): return int(self._get("mstate", i32))
    def n_multi(self): return int(self._get("n_multi", i32))
    def kbt(self): return float(self._get("kbt", f64))
    def beta(self): return float(self._get("beta", f64))
    def sync(self): self._call("sync")
    def set_kbt(self, kbt): self._call("set_kbt", float(kbt), argtypes=(f64,))
    def skip_curand_clock(self, n_skip): self._call("skip_curand_clock", int(n_skip), argtypes=(i64,))
    def init_sixclock_order(self): self._call("init_sixclock_order")
    def update_metropolis(self): self._call("update_metropolis")
    def update_metropolis_n(self, n): self._call("update_metropolis_n", int(n), argtypes=(i32,))

    def set_sample_offset(self, first_sample):
        """this handle's samples are samples first_sample .. first_sample + n_multi - 1 of the job (a batch split across
        GPUs, one handle per rank, no exchange); call right after construction"""
        self._call("set_sample_offset", int(first_sample), argtypes=(i32,))
        return self

    @classmethod
    def distributed(cls, nx, ny, kbt, mstate, n_multi, iseed, group=None, variant=0):
        """the n_multi samples of the job shared out over the ranks of a torch.distributed group (independent samples:
        no collective on the data path); calc_energy_all / calc_magne_all gather the per-sample values of the whole job"""
        import torch.distributed as dist
        from .clock_gpu_multi_m import split_samples
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        lo, hi = split_samples(n_multi, rank, world)
        if hi == lo:
            raise ValueError(f"rank {rank}: no sample to run ({n_multi} samples on {world} ranks)")
        self = cls(nx, ny, kbt, mstate, hi - lo, iseed, variant)
        self.set_sample_offset(lo)
        self._dist = (dist, group)
        return self

    def _gather(self, local):
        dist, group = self._dist
        parts = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, [float(x) for x in np.atleast_1d(local)], group=group)
        return np.array([x for p in parts for x in p], dtype=np.float64)

    def calc_energy_all(self):
        return self._gather(self.calc_energy())

    def calc_magne_all(self):
        return self._gather(self.calc_magne())

    def update_with_rnds(self, rnds):
        r = np.ascontiguousarray(rnds, dtype=np.float64)
        if r.size != 2 * self.nall() * self.n_multi():
            raise ValueError("rnds must hold 2 * nall (x n_multi) uniforms, order rnds(2, nx, ny)")
        self._call("update_with_rnds", r.ctypes.data_as(P), argtypes=(P,))

    def _vec(self, name):
        out = np.empty(self.n_multi(), dtype=np.float64)
        self._call(name, out.ctypes.data_as(P), argtypes=(P,))
        return out

    def calc_energy(self): return self._vec("calc_energy")
    def calc_magne(self): return self._vec("calc_magne")

    def histograms(self):
        q, n = self.mstate(), self.n_multi()
        h = np.zeros((n, q), dtype=np.int64)
        br = np.zeros((n, q), dtype=np.int64)
        bu = np.zeros((n, q), dtype=np.int64)
        self._call("get_histograms", h.ctypes.data_as(P), br.ctypes.data_as(P), bu.ctypes.data_as(P), argtypes=(P, P, P))
        return h, br, bu

    def get_sixclock(self):
        """sixclock(nx, ny) int32 in Fortran order, one flat array per sample: shape (n_multi, nall)"""
        out = np.empty((self.n_multi(), self.nall()), dtype=np.int32)
        self._call("get_sixclock", out.ctypes.data_as(P), argtypes=(P,))
        return out

    def set_sixclock(self, a):
        s = np.ascontiguousarray(a, dtype=np.int32)
        if s.size != self.n_multi() * self.nall():
            raise ValueError("sixclock must hold nall (x n_multi) states")
        self._call("set_sixclock", s.ctypes.data_as(P), argtypes=(P,))

    def get_dual(self):
        """sixclock_even, sixclock_odd (nx/2, ny) int32 in Fortran order: shapes (n_multi, nall/2)"""
        ev = np.empty((self.n_multi(), self.nall() // 2), dtype=np.int32)
        od = np.empty_like(ev)
        self._call("get_dual", ev.ctypes.data_as(P), od.ctypes.data_as(P), argtypes=(P, P))
        return ev, od

    def set_dual(self, even, odd):
        ev = np.ascontiguousarray(even, dtype=np.int32)
        od = np.ascontiguousarray(odd, dtype=np.int32)
        if ev.size != self.n_multi() * self.nall() // 2 or od.size != ev.size:
            raise ValueError("even / odd must hold nall/2 (x n_multi) states each")
        self._call("set_dual", ev.ctypes.data_as(P), od.ctypes.data_as(P), argtypes=(P, P))

    def states_to_prob(self):
        q = self.mstate()
        out = np.empty(q ** 6, dtype=np.float64)
        self._call("get_states_to_prob", out.ctypes.data_as(P), argtypes=(P,))
        return out

    def energy_table(self):
        q = self.mstate()
        out = np.empty(q ** 3, dtype=np.float64)
        self._call("get_energy_table", out.ctypes.data_as(P), argtypes=(P,))
        return out

    def set_timing(self, on): self._call("set_timing", 1 if on else 0, argtypes=(i32,))

    def get_timing(self):
        n, ms = i64(0), f64(0.0)
        self._call("get_timing", C.byref(n), C.byref(ms), argtypes=(C.POINTER(i64), C.POINTER(f64)))
        return int(n.value), float(ms.value)
