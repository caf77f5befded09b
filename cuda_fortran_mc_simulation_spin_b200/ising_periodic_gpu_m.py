"""Ising 2D / 3D with true periodic boundaries (torus) -- host mirror over ``b200mc_ising_torus_*``.

Not a reference module: ``type(ising3d_gpu)`` / ``type(ising2d_gpu)`` are helical and valid for odd ``nx`` only
(src/ising3d_gpu_m.f90:60-62,196), so L = 1024^3 -- the size BASELINE.json names -- needs this boundary
condition.  Same procedure names, update rule, tables, value conventions (3D 0 / 1, 2D -1 / +1) and observables as
those types; host arrays are ``s[z][y][x]`` without halo cells.  ``nx % 32 == 0``, ``ny`` and ``nz`` even.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import P, PP, f64, i32, i64

METROPOLIS, HEATBATH = 0, 1
ising_periodic_gpu_stat = 0


class ising_periodic_gpu:
    _pfx = "b200mc_ising_torus"

    def __init__(self):
        self._h = C.c_void_p(None)
        self._dims = None

    def _f(self, name, restype, *argtypes):
        return _lib.fn(f"{self._pfx}_{name}", restype, *argtypes)

    def _call(self, name, *args, argtypes=()):
        _lib.check(self._f(name, C.c_int, P, *argtypes)(self._h, *args))

    def __del__(self):
        try:
            if self._h:
                self._f("destroy", C.c_int, P)(self._h)
                self._h = C.c_void_p(None)
        except Exception:
            pass

    def init(self, nx, ny, nz, kbt, iseed):
        """init(nx, ny, nz, kbt, iseed) as init_ising3d_gpu (src/ising3d_gpu_m.f90:50-71); nz = 0 -> the 2D model"""
        if self._h:
            self._f("destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        ndim = 3 if int(nz) > 0 else 2
        f = _lib.fn("b200mc_ising_torus_create", C.c_int, PP, i32, i64, i64, i64, f64, i32)
        _lib.check(f(C.byref(self._h), ndim, int(nx), int(ny), int(nz), float(kbt), int(iseed)))
        self._dims = (int(nz), int(ny), int(nx)) if ndim == 3 else (int(ny), int(nx))
        return self

    def init_distributed(self, nx, ny, nz, kbt, iseed, group=None):
        """the global nx x ny x nz lattice, one slab of planes per rank of an initialised torch.distributed job (one process per
        GPU): rank 0 draws the NCCL id of the library's own communicator, the job's process group only broadcasts those 128 bytes"""
        import torch.distributed as dist
        from ._ising_base import unique_id

        if self._h:
            self._f("destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        rank, nranks = dist.get_rank(group), dist.get_world_size(group)
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        f = _lib.fn("b200mc_ising_torus_create_slab", C.c_int, PP, i32, i64, i64, i64, f64, i32, i32, i32, C.c_char_p)
        _lib.check(f(C.byref(self._h), 3, int(nx), int(ny), int(nz), float(kbt), int(iseed), rank, nranks, bytes(box[0])))
        self._group, self._dist = group, dist
        self._dims = (int(nz), int(ny), int(nx))
        self._p2p = False
        if os.environ.get("B200MC_SLAB_TRANSPORT", "p2p") != "nccl":
            self._connect_p2p(dist, group, rank, nranks)
        return self

    def _connect_p2p(self, dist, group, rank, nranks):
        """direct transport: exchange the CUDA IPC handles and map the two neighbours' arrays; every rank falls back to the NCCL
        transport if any rank cannot export or map"""
        buf = C.create_string_buffer(192)
        ok = 1
        try:
            self._call("p2p_handles", buf, argtypes=(C.c_char_p,))
        except _lib.B200MCError:
            ok = 0
        allh = [None] * nranks
        dist.all_gather_object(allh, (ok, buf.raw), group=group)
        if not all(o for o, _ in allh):
            return
        prev, nxt = allh[(rank - 1) % nranks][1], allh[(rank + 1) % nranks][1]
        rc = self._f("p2p_connect", C.c_int, P, C.c_char_p, C.c_char_p)(self._h, prev, nxt)
        oks = [None] * nranks
        dist.all_gather_object(oks, rc == 0, group=group)
        if not all(oks):
            raise _lib.B200MCError("CUDA IPC mapping of the neighbour slabs failed on some rank; rerun with B200MC_SLAB_TRANSPORT=nccl")
        self._p2p = True

    def rank_info(self):
        """(rank, nranks, first owned plane, owned planes)"""
        r, n, z0, nzl = C.c_int32(0), C.c_int32(1), C.c_int64(0), C.c_int64(0)
        self._call("rank_info", C.byref(r), C.byref(n), C.byref(z0), C.byref(nzl), argtypes=(P, P, P, P))
        return int(r.value), int(n.value), int(z0.value), int(nzl.value)

    def _nlocal(self):
        _, n, _, nzl = self.rank_info()
        return self.nall() // n if n > 1 else self.nall()

    def set_allup_spin(self): self._call("set_allup_spin")
    def set_random_spin(self): self._call("set_random_spin")
    def set_kbt(self, kbt): self._call("set_kbt", float(kbt), argtypes=(f64,))
    def set_beta(self, beta): self._call("set_beta", float(beta), argtypes=(f64,))
    def set_method(self, method): self._call("set_method", int(method), argtypes=(i32,))
    def skip_curand(self, n_skip): self._call("skip_curand", int(n_skip), argtypes=(i64,))
    def update(self): self._call("update")
    def update_n(self, n_sweeps): self._call("update_n", int(n_sweeps), argtypes=(i32,))

    def update_with_randoms(self, randoms):
        r = np.ascontiguousarray(randoms, dtype=np.float64).ravel()
        assert r.size == self.nall()
        self._call("update_with_randoms", r.ctypes.data_as(P), argtypes=(P,))

    def nall(self): return int(self._f("nall", i64, P)(self._h))
    def beta(self): return float(self._f("beta", f64, P)(self._h))
    def kbt(self): return 1.0 / self.beta()

    def table(self):
        """w[s, S]: acceptance probability of a site with spin s (0 / 1) and S up neighbours, as the reference builds it"""
        out = np.empty(16, dtype=np.float64)
        self._call("get_table", out.ctypes.data_as(P), argtypes=(P,))
        return out.reshape(2, 8)

    def spins_local(self):
        """the planes this rank owns (the whole lattice on a single-GPU handle)"""
        out = np.empty(self._nlocal(), dtype=np.int32)
        self._call("get_spins", out.ctypes.data_as(P), argtypes=(P,))
        return out

    def spins(self):
        """the whole lattice, s[z][y][x] flattened (slab mode: gathered over the ranks, every rank gets it)"""
        loc = self.spins_local()
        rank, n, _, _ = self.rank_info()
        if n == 1:
            return loc
        import torch
        t = torch.from_numpy(loc).cuda()
        parts = [torch.empty_like(t) for _ in range(n)]
        self._dist.all_gather(parts, t, group=self._group)
        return torch.cat(parts).cpu().numpy()

    def set_spins(self, spins):
        s = np.ascontiguousarray(spins, dtype=np.int32).ravel()
        assert s.size == self.nall()
        rank, n, _, _ = self.rank_info()
        if n > 1:
            nl = self._nlocal()
            s = np.ascontiguousarray(s[rank * nl:(rank + 1) * nl])
        self._call("set_spins", s.ctypes.data_as(P), argtypes=(P,))

    def measure(self):
        e, m = C.c_int64(0), C.c_int64(0)
        self._call("measure", C.byref(e), C.byref(m), argtypes=(P, P))
        return int(e.value), int(m.value)

    def calc_energy_sum(self):
        e = C.c_int64(0)
        self._call("calc_energy_sum", C.byref(e), argtypes=(P,))
        return int(e.value)

    def calc_magne_sum(self):
        m = C.c_int64(0)
        self._call("calc_magne_sum", C.byref(m), argtypes=(P,))
        return int(m.value)

    def sync(self): self._call("sync")
    def set_timing(self, on=True): self._call("set_timing", int(bool(on)), argtypes=(i32,))

    def get_timing(self):
        n, ms = C.c_int64(0), C.c_double(0.0)
        self._call("get_timing", C.byref(n), C.byref(ms), argtypes=(P, P))
        return int(n.value), float(ms.value)
