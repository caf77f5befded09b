"""Host mirror of ``module xy2d_periodic_gpu_m`` (src/xy2d_periodic_gpu_m.f90): ``type(xy2d_gpu)``
(:14-59) with the reference's procedure names, over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import P, PP, f64, i32, i64

xy2d_gpu_stat = 0  # mirrors `integer(int32), public, protected :: xy2d_gpu_stat` (:9)


class xy2d_gpu:
    def __init__(self):
        self._h = C.c_void_p(None)

    def _call(self, name, *args, argtypes=()):
        f = _lib.fn(f"b200mc_xy2d_{name}", C.c_int, P, *argtypes)
        _lib.check(f(self._h, *args))

    def _get(self, name, restype):
        return _lib.fn(f"b200mc_xy2d_{name}", restype, P)(self._h)

    def _dbl(self, name):
        r = C.c_double(0.0)
        self._call(name, C.byref(r), argtypes=(C.POINTER(C.c_double),))
        return float(r.value)

    def __del__(self):
        try:
            if self._h:
                _lib.fn("b200mc_xy2d_destroy", C.c_int, P)(self._h)
                self._h = C.c_void_p(None)
        except Exception:
            pass

    def init(self, nx, ny, kbt, iseed):
        """init_xy2d_gpu(this, nx, ny, kbt, iseed), :61-78"""
        if self._h:
            _lib.fn("b200mc_xy2d_destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        f = _lib.fn("b200mc_xy2d_create", C.c_int, PP, i64, i64, f64, i32)
        _lib.check(f(C.byref(self._h), int(nx), int(ny), float(kbt), int(iseed)))
        return self

    def init_distributed(self, nx, ny, kbt, iseed, group=None):
        """slabs along y over the ranks of an initialised torch.distributed job (one process per GPU): this rank holds
        ny / world rows; halo rows go between neighbouring ranks after every colour pass, E / Mx / My are all-reduced, and
        every site draws the random numbers of the one-GPU run.  ny() is then the local row count, nall() the lattice."""
        import torch.distributed as dist
        from ._ising_base import unique_id

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if self._h:
            _lib.fn("b200mc_xy2d_destroy", C.c_int, P)(self._h)
            self._h = C.c_void_p(None)
        f = _lib.fn("b200mc_xy2d_create_slab", C.c_int, PP, i64, i64, f64, i32, i32, i32, C.c_char_p)
        _lib.check(f(C.byref(self._h), int(nx), int(ny), float(kbt), int(iseed), int(rank), int(world), bytes(box[0])))
        self._dist = (dist, group)
        return self

    def angles_all(self):
        """slab mode: the angles of the whole lattice [ny_global][nx], gathered from the ranks' rows"""
        dist, group = self._dist
        parts = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, self.angles(), group=group)
        return np.concatenate(parts, axis=0)

    def skip_curand(self, n_skip): self._call("skip_curand", int(n_skip), argtypes=(i64,))
    def set_allup_spin(self): self._call("set_allup_spin")
    def set_random_spin(self): self._call("set_random_spin")
    def set_random_small_spin(self, near_magne): self._call("set_random_small_spin", float(near_magne), argtypes=(f64,))  # :158-175
    def set_random_near_spin(self, near_magne, diff_parcent): self._call("set_random_near_spin", float(near_magne), float(diff_parcent), argtypes=(f64, f64))  # :179-196
    def set_finite_magne_spin(self, init_magne): self._call("set_finite_magne_spin", float(init_magne), argtypes=(f64,))  # :126-154
    def metropolis_by_field(self, hx, hy): self._call("metropolis_by_field", float(hx), float(hy), argtypes=(f64, f64))  # one launch of :198-216
    def set_kbt(self, kbt): self._call("set_kbt", float(kbt), argtypes=(f64,))
    def set_beta(self, beta): self._call("set_beta", float(beta), argtypes=(f64,))
    def update(self): self._call("update")
    def update_n(self, n): self._call("update_n", int(n), argtypes=(i32,))

    def update_with_randoms(self, randoms, candidates):
        """one Metropolis MCS on the caller's uniforms: randoms(nx, ny), candidates(nx, ny) as the reference fills them (:355-356)"""
        r = np.ascontiguousarray(randoms, dtype=np.float64)
        c = np.ascontiguousarray(candidates, dtype=np.float64)
        if r.size != self.nall() or c.size != self.nall():
            raise ValueError("randoms / candidates must hold nall uniforms each")
        self._call("update_with_randoms", r.ctypes.data_as(P), c.ctypes.data_as(P), argtypes=(P, P))

    def update_over_relaxation(self, n_steps): self._call("update_over_relaxation", int(n_steps), argtypes=(i32,))
    def set_initial_magne_autocorrelation_state(self): self._call("set_initial_magne_autocorrelation_state")
    def rotate_summation_magne_toward_xaxis(self): self._call("rotate_summation_magne_toward_xaxis", 0, argtypes=(i32,))
    def rotate_summation_magne_and_autocorrelation_toward_xaxis(self): self._call("rotate_summation_magne_toward_xaxis", 1, argtypes=(i32,))
    def nx(self): return int(self._get("nx", i64))
    def ny(self): return int(self._get("ny", i64))
    def nall(self): return int(self._get("nall", i64))
    def kbt(self): return float(self._get("kbt", f64))
    def beta(self): return float(self._get("beta", f64))
    def sync(self): self._call("sync")
    def calc_energy_sum(self): return self._dbl("calc_energy_sum")
    def calc_magne_sum(self): return self._dbl("calc_magne_sum")
    def calc_magne_y_sum(self): return self._dbl("calc_magne_y_sum")
    def calc_autocorrelation_sum(self): return self._dbl("calc_autocorrelation_sum")
    def calc_correlation_sum(self): return self._dbl("calc_correlation_sum")

    def measure(self):
        e, mx, my = C.c_double(0), C.c_double(0), C.c_double(0)
        pd = C.POINTER(C.c_double)
        self._call("measure", C.byref(e), C.byref(mx), C.byref(my), argtypes=(pd, pd, pd))
        return float(e.value), float(mx.value), float(my.value)

    def spins(self):
        """real64 (cos, sin) in the reference layout spins(0:nx+1, 0:ny+1, 1:2): array [2][ny+2][nx+2]"""
        out = np.empty((2, self.ny() + 2, self.nx() + 2), dtype=np.float64)
        self._call("get_spins", out.ctypes.data_as(P), argtypes=(P,))
        return out

    def angles(self):
        """native state: fp32 angles in turns, [ny][nx]"""
        out = np.empty((self.ny(), self.nx()), dtype=np.float32)
        self._call("get_angles", out.ctypes.data_as(P), argtypes=(P,))
        return out

    def set_angles(self, turns):
        t = np.ascontiguousarray(turns, dtype=np.float32)
        if t.size != self.nall():
            raise ValueError("angles must be [ny][nx]")
        self._call("set_angles", t.ctypes.data_as(P), argtypes=(P,))
