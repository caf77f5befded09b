"""Shared host-side plumbing of the two helical Ising mirrors (ctypes over the C ABI)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import P, PP, f64, i32, i64

METROPOLIS = 0
HEATBATH = 1


def unique_id() -> bytes:
    """NCCL unique id for a slab-decomposed job (call on rank 0, broadcast to the others)"""
    buf = C.create_string_buffer(128)
    _lib.check(_lib.fn("b200mc_dist_unique_id", C.c_int, C.c_char_p)(buf))
    return buf.raw


def slab_geometry(nx, ny, nz, rank, nranks):
    """host-only: {Nc, L, H, p0, Lloc, ptail} of the slab `rank` of `nranks` would own (nz = 0: 2D)"""
    out = (C.c_int64 * 6)()
    _lib.check(_lib.fn("b200mc_ring_slab_geometry", C.c_int, i64, i64, i64, i32, i32, C.POINTER(C.c_int64))(
        int(nx), int(ny), int(nz), int(rank), int(nranks), out))
    return dict(zip(("Nc", "L", "H", "p0", "Lloc", "ptail"), [int(x) for x in out]))


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

    # -- bit-packed storage (one bit per site, Metropolis, one GPU; csrc/ising_bits.cu) ----
    def _init_packed(self, dims, kbt, iseed):
        """the same type over the bit-packed handle: same procedures and host layouts, its own random stream"""
        if self._h:
            self._f("destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        if not getattr(self, "_packed", False):
            self._pfx = self._pfx + "p"      # b200mc_ising3dp_* / b200mc_ising2dp_*
            self._packed = True
        f = self._f("create", C.c_int, PP, *([i64] * len(dims)), f64, i32)
        _lib.check(f(C.byref(self._h), *[int(d) for d in dims], float(kbt), int(iseed)))
        return self

    def _init_packed_distributed(self, dims, kbt, iseed, group=None):
        """bit-packed storage in slab mode, driven by an initialised torch.distributed job (halos through NCCL send/recv)"""
        import torch.distributed as dist

        rank, nranks = dist.get_rank(group), dist.get_world_size(group)
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if self._h:
            self._f("destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        if not getattr(self, "_packed", False):
            self._pfx = self._pfx + "p"
            self._packed = True
        f = self._f("create_slab", C.c_int, PP, *([i64] * len(dims)), f64, i32, i32, i32, C.c_char_p)
        _lib.check(f(C.byref(self._h), *[int(d) for d in dims], float(kbt), int(iseed), int(rank), int(nranks), bytes(box[0])))
        self._group = group
        self._rank, self._nranks = int(rank), int(nranks)
        return self

    # -- slab decomposition over ranks (one process per GPU) ----
    def _init_slab(self, dims, kbt, iseed, rank, nranks, nccl_id):
        if self._h:
            self._f("destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        if len(nccl_id) != 128:
            raise ValueError("nccl_id must be the 128 bytes of b200mc_dist_unique_id")
        f = self._f("create_slab", C.c_int, PP, *([i64] * len(dims)), f64, i32, i32, i32, C.c_char_p)
        _lib.check(f(C.byref(self._h), *[int(d) for d in dims], float(kbt), int(iseed), int(rank), int(nranks),
                     bytes(nccl_id)))
        self._rank, self._nranks = int(rank), int(nranks)
        return self

    def _init_torch_distributed(self, dims, kbt, iseed, group=None):
        """slab mode driven by an initialised torch.distributed job: rank 0 draws the NCCL id of the
        library's own communicator, the job's process group only broadcasts those 128 bytes"""
        import torch.distributed as dist

        rank, nranks = dist.get_rank(group), dist.get_world_size(group)
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._group = group
        self._init_slab(dims, kbt, iseed, rank, nranks, box[0])
        if os.environ.get("B200MC_SLAB_TRANSPORT", "p2p") != "nccl":
            self._connect_p2p(dist, group, rank, nranks)
        return self

    def _connect_p2p(self, dist, group, rank, nranks):
        """direct NVLink transport: exchange the CUDA IPC handles and map the two neighbours' arrays;
        falls back to the NCCL transport (on every rank) if any rank cannot map its neighbours"""
        buf = C.create_string_buffer(192)
        ok = 1
        try:
            self._call("p2p_handles", buf, argtypes=(C.c_char_p,))
        except _lib.B200MCError:
            ok = 0
        allh = [None] * nranks
        dist.all_gather_object(allh, (ok, buf.raw), group=group)
        if not all(o for o, _ in allh):
            self._p2p = False
            return
        prev, nxt = allh[(rank - 1) % nranks][1], allh[(rank + 1) % nranks][1]
        # mapping must succeed everywhere or nowhere: connect first, then agree
        f = self._f("p2p_connect", C.c_int, P, C.c_char_p, C.c_char_p)
        rc = f(self._h, prev, nxt)
        oks = [None] * nranks
        dist.all_gather_object(oks, rc == 0, group=group)
        if not all(oks):
            raise _lib.B200MCError("CUDA IPC mapping of the neighbour slabs failed on some rank; "
                                   "rerun with B200MC_SLAB_TRANSPORT=nccl")
        self._p2p = True
        # observables: every rank's mailbox mapped on every rank (all or none, like the slabs)
        self._sums_p2p = False
        if nranks <= 16:
            blob = b"".join(h for _, h in allh)
            rc = self._f("p2p_connect_sums", C.c_int, P, C.c_char_p)(self._h, blob)
            oks = [None] * nranks
            dist.all_gather_object(oks, rc == 0, group=group)
            if not all(oks):
                raise _lib.B200MCError("CUDA IPC mapping of the ranks' flag buffers failed on some rank; "
                                       "rerun with B200MC_SLAB_TRANSPORT=nccl")
            self._sums_p2p = True

    def rank_info(self):
        r, n = C.c_int32(0), C.c_int32(1)
        self._call("rank_info", C.byref(r), C.byref(n), argtypes=(C.POINTER(C.c_int32), C.POINTER(C.c_int32)))
        return int(r.value), int(n.value)

    def spins_local(self):
        """spins() of this rank only: sites owned by other ranks read INT32_MIN"""
        out = np.empty(self.nall() + 2 * self._halo(), dtype=np.int32)
        self._call("get_spins", out.ctypes.data_as(P), argtypes=(P,))
        return out

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
        out = self.spins_local()
        if getattr(self, "_nranks", 1) > 1:
            import torch
            import torch.distributed as dist

            t = torch.from_numpy(out)
            if dist.get_backend(getattr(self, "_group", None)) == "nccl":
                t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=getattr(self, "_group", None))
            out = t.cpu().numpy()
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

    def run_relaxation(self, mcs):
        """mcs x [update; calc_magne_sum; calc_energy_sum] on the device; returns (E, M) int64 arrays of shape
        (mcs,) -- or (n_multi, mcs) for a batch of samples"""
        n = self.n_multi()
        e = np.empty((n, int(mcs)), dtype=np.int64)
        m = np.empty((n, int(mcs)), dtype=np.int64)
        self._call("run_relaxation", int(mcs), e.ctypes.data_as(P), m.ctypes.data_as(P), argtypes=(i32, P, P))
        return (e[0], m[0]) if n == 1 else (e, m)

    def run_relaxation_stats(self, mcs, tot_sample, random_start=False):
        """the drivers' whole measurement on the device (app/ising3d_gpu_relaxation.f90:37-55): returns an (mcs, 8) array
        [num_sample, <m>, <e>, <m^2>, <e^2>, var m, var e, cov(m, e)] per MCS, m and e per site"""
        out = np.empty((int(mcs), 8), dtype=np.float64)
        self._call("run_relaxation_stats", int(mcs), int(tot_sample), 1 if random_start else 0, out.ctypes.data_as(P),
                   argtypes=(i32, i32, i32, P))
        return out

    def format_relaxation_table(self, stats):
        """the rows the drivers print (:49-55): nall, num_sample, i, <m>, <e>, <m^2>, <e^2>, N var m, N var e, N cov"""
        f = _lib.fn("b200mc_format_relaxation_row", C.c_int, i64, i32, P, C.c_char_p, i32)
        buf = C.create_string_buffer(512)
        rows = []
        st = np.ascontiguousarray(stats, dtype=np.float64)
        for i in range(st.shape[0]):
            _lib.check(f(self.nall(), i + 1, st[i].ctypes.data_as(P), buf, 512))
            rows.append(buf.value.decode())
        return rows

    # -- batch of independent samples (one launch per colour pass for all of them) --
    def init_multi(self, *dims_kbt_iseed_nmulti):
        """init(nx, ny[, nz], kbt, iseed) for n_multi samples: init_multi(nx, ny[, nz], kbt, iseed, n_multi)"""
        *dims, kbt, iseed, n_multi = dims_kbt_iseed_nmulti
        if self._h:
            self._f("destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        f = self._f("create_multi", C.c_int, PP, *([i64] * len(dims)), f64, i32, i32)
        _lib.check(f(C.byref(self._h), *[int(d) for d in dims], float(kbt), int(iseed), int(n_multi)))
        return self

    def n_multi(self):
        return int(self._f("n_multi", i32, P)(self._h))

    def measure_multi(self):
        n = self.n_multi()
        e = np.empty(n, dtype=np.int64)
        m = np.empty(n, dtype=np.int64)
        self._call("measure_multi", e.ctypes.data_as(P), m.ctypes.data_as(P), argtypes=(P, P))
        return e, m

    def spins_multi(self, sample):
        out = np.empty(self.nall() + 2 * self._halo(), dtype=np.int32)
        self._call("get_spins_multi", int(sample), out.ctypes.data_as(P), argtypes=(i32, P))
        return out

    def set_spins_multi(self, sample, spins):
        s = np.ascontiguousarray(spins, dtype=np.int32)
        if s.size != self.nall() + 2 * self._halo():
            raise ValueError("spins must use the reference layout, halo cells included")
        self._call("set_spins_multi", int(sample), s.ctypes.data_as(P), argtypes=(i32, P))

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
