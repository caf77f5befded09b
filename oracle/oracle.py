"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

ctypes binding + thin numpy classes over ``oracle/liboracle.so`` (built from
oracle.c / rng_contract.c by ``make -C oracle``).  The classes mirror the
reference's module types (``ising2d_gpu``, ``ising3d_gpu``, ``clock_gpu``,
``xy2d_gpu`` and the module-procedure API of ``clock_tableall_gpu_m``) with the
reference's array layouts (halo cells included), so a parity test can compare
``spins()`` of the CUDA build to ``spins()`` of the oracle element for element.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.

PARITY UNPINNED: the reference holds no golden vectors and cannot be built in
this image (see oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "rng_contract.c", "philox.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


i64, i32, u32, u64, f64 = C.c_int64, C.c_int32, C.c_uint32, C.c_uint64, C.c_double
P = C.c_void_p


def _declare(L):
    def d(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    d("orc_max_threads", C.c_int)
    d("orc_set_threads", None, C.c_int)
    d("orc_philox", None, P, P, P)
    d("orc_ising2d_exparr", None, f64, P)
    d("orc_ising2d_norishiro", None, i64, i64, P)
    d("orc_ising2d_set_allup", None, i64, i64, P)
    d("orc_ising2d_set_random", None, i64, i64, P, P)
    d("orc_ising2d_update", None, i64, i64, P, P, P)
    d("orc_ising2d_energy", i64, i64, i64, P)
    d("orc_ising2d_magne", i64, i64, i64, P)
    d("orc_ising3d_tables", None, f64, P, P)
    d("orc_ising3d_norishiro", None, i64, i64, i64, P)
    d("orc_ising3d_set_allup", None, i64, i64, i64, P)
    d("orc_ising3d_set_random", None, i64, i64, i64, P, P)
    d("orc_ising3d_update", None, i64, i64, i64, P, P, P)
    d("orc_ising3d_energy", i64, i64, i64, i64, P, P)
    d("orc_ising3d_magne", i64, i64, i64, i64, P)
    d("orc_kahan_add_data", None, P, f64, f64)
    d("orc_kahan_results", None, P, i64, P)
    d("orc_heatbath_table", None, f64, C.c_int, P)
    d("orc_ising2d_update_heatbath", None, i64, i64, P, P, P)
    d("orc_ising3d_update_heatbath", None, i64, i64, i64, P, P, P)
    d("orc_clock_tables", None, i32, f64, P, P, P)
    d("orc_clock_norishiro", None, i64, i64, P)
    d("orc_clock_set_random", None, i64, i64, i32, P, P)
    d("orc_clock_update", None, i64, i64, i32, P, P, P, P, C.c_int)
    d("orc_clock_energy", f64, i64, i64, i32, P, P)
    d("orc_clock_magne", f64, i64, i64, i32, P, P)
    d("orc_clock_histograms", None, i64, i64, i32, P, P, P)
    d("orc_tableall_tables", None, i32, f64, P, P, P)
    d("orc_tableall_update", None, i64, i64, i32, P, P, P)
    d("orc_clock_simple_prob", None, i32, f64, P)
    d("orc_tableall_magne", f64, i64, i64, i32, P, P)
    d("orc_tableall_energy", f64, i64, i64, i32, P, P)
    d("orc_tableall_histograms", None, i64, i64, i32, P, P, P)
    d("orc_dual_update", None, i64, i64, i32, P, P, P, P)
    d("orc_dual_energy", f64, i64, i64, i32, P, P, P)
    d("orc_xy_norishiro", None, i64, i64, P)
    d("orc_xy_set_allup", None, i64, i64, P)
    d("orc_xy_set_random", None, i64, i64, P, P, C.c_int)
    d("orc_xy_update", None, i64, i64, P, f64, P, P)
    d("orc_xy_over_relaxation", None, i64, i64, P, i32)
    d("orc_xy_energy", f64, i64, i64, P)
    d("orc_xy_magne", f64, i64, i64, P, C.c_int)
    d("orc_xy_autocorrelation", f64, i64, i64, P, P)
    d("orc_xy_correlation", f64, i64, i64, P)
    d("orc_xy_rotate", None, i64, i64, P, f64)
    d("orc_xy_metropolis_by_field", None, i64, i64, P, P, P, f64, f64)
    d("orc_isingp_update", None, C.c_int, i64, i64, i64, P, P, P, C.c_int)
    d("orc_isingp_energy_magne", None, C.c_int, i64, i64, i64, P, P, P)
    d("orc_isingp_uniforms", None, u32, u64, i64, i64, i64, P)
    d("orc_isingp_init_uniforms", None, u32, u64, i64, i64, i64, P)
    d("orc_ring_fold_len", i64, i64)
    d("orc_ising_uniforms", None, u32, u64, i64, P)
    d("orc_ising_uniforms_fast", None, u32, u64, i64, P)
    d("orc_ising_uniforms_rep", None, u32, u64, u32, i64, P)
    d("orc_ring_init_uniforms_rep", None, u32, u64, u32, i64, P)
    d("orc_ring_init_uniforms", None, u32, u64, i64, P)
    d("orc_clock_uniforms", None, u32, u64, i32, i64, i32, P, P)
    d("orc_xy_uniforms", None, u32, u64, i64, i64, P, P)
    d("orc_xy_init_uniforms", None, u32, u64, i64, i64, P)
    d("orc_torus_uniforms", None, u32, u64, i32, i64, i64, i32, P)
    d("orc_isingbits_uniforms", C.c_int, u32, u64, i64, C.c_int, P)
    d("orc_xyh_norishiro", None, i64, i64, P)
    d("orc_xyh_set_allup", None, i64, i64, P)
    d("orc_xyh_set_random", None, i64, i64, P, P)
    d("orc_xyh_update", None, i64, i64, P, f64, P, P)
    d("orc_xyh_over_relaxation", None, i64, i64, P, i32)
    d("orc_xyh_energy", f64, i64, i64, P)
    d("orc_xyh_magne", f64, i64, i64, P)
    d("orc_xyh_uniforms", None, u32, u64, i64, P, P)
    d("orc_xyh_init_uniforms", None, u32, u64, i64, P)


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(P)


def philox(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(_p(c), _p(k), _p(out))
    return out


def max_threads() -> int:
    return lib().orc_max_threads()


def set_threads(n: int) -> None:
    lib().orc_set_threads(n)


# --------------------------------------------------------------------------
# RNG contract arrays (what the reference would have read from cuRAND)
# --------------------------------------------------------------------------
def ising_uniforms(seed: int, draw: int, n_sites: int) -> np.ndarray:
    out = np.empty(n_sites, dtype=np.float64)
    lib().orc_ising_uniforms(seed & 0xFFFFFFFF, draw, n_sites, _p(out))
    return out


def ising_uniforms_fast(seed: int, draw: int, n_sites: int, out=None) -> np.ndarray:
    if out is None:
        out = np.empty(n_sites, dtype=np.float64)
    lib().orc_ising_uniforms_fast(seed & 0xFFFFFFFF, draw, n_sites, _p(out))
    return out


def ising_uniforms_rep(seed: int, draw: int, rep: int, n_sites: int) -> np.ndarray:
    """accept uniforms of sample `rep` of a batch (b200mc_ising*_create_multi)"""
    out = np.empty(n_sites, dtype=np.float64)
    lib().orc_ising_uniforms_rep(seed & 0xFFFFFFFF, draw, rep, n_sites, _p(out))
    return out


def ring_init_uniforms_rep(seed: int, draw: int, rep: int, n_sites: int) -> np.ndarray:
    out = np.empty(n_sites, dtype=np.float64)
    lib().orc_ring_init_uniforms_rep(seed & 0xFFFFFFFF, draw, rep, n_sites, _p(out))
    return out


def ring_init_uniforms(seed: int, draw: int, n_sites: int) -> np.ndarray:
    out = np.empty(n_sites, dtype=np.float64)
    lib().orc_ring_init_uniforms(seed & 0xFFFFFFFF, draw, n_sites, _p(out))
    return out


def clock_uniforms(seed: int, draw: int, replica: int, n_sites: int, q: int = 6):
    r = np.empty(n_sites, dtype=np.float64)
    p = np.empty(n_sites, dtype=np.float64)
    lib().orc_clock_uniforms(seed & 0xFFFFFFFF, draw, replica, n_sites, q, _p(r), _p(p))
    return r, p


def xy_uniforms(seed: int, draw: int, nx: int, ny: int):
    r = np.empty(nx * ny, dtype=np.float64)
    c = np.empty(nx * ny, dtype=np.float64)
    lib().orc_xy_uniforms(seed & 0xFFFFFFFF, draw, nx, ny, _p(r), _p(c))
    return r, c


def torus_uniforms(seed: int, draw: int, replica: int, nx: int, ny: int, q: int = 6) -> np.ndarray:
    """rnds(2, nx, ny) of update_metropolis (src/clock/clock_tableall_gpu_m.f90:95), flat"""
    out = np.empty(2 * nx * ny, dtype=np.float64)
    lib().orc_torus_uniforms(seed & 0xFFFFFFFF, draw, replica, nx, ny, q, _p(out))
    return out


def isingbits_uniforms(seed: int, draw: int, n_sites: int, init: bool = False) -> np.ndarray:
    """accept uniforms (init: set_random_spin uniforms) of the bit-packed Ising handles, reference index order"""
    out = np.empty(n_sites, dtype=np.float64)
    if lib().orc_isingbits_uniforms(seed & 0xFFFFFFFF, draw, n_sites, 1 if init else 0, _p(out)):
        raise ValueError("bit-packed Ising: n_sites / 2 must be a multiple of 128")
    return out


def xyh_uniforms(seed: int, draw: int, n_sites: int):
    r = np.empty(n_sites, dtype=np.float64)
    c = np.empty(n_sites, dtype=np.float64)
    lib().orc_xyh_uniforms(seed & 0xFFFFFFFF, draw, n_sites, _p(r), _p(c))
    return r, c


def xyh_init_uniforms(seed: int, draw: int, n_sites: int):
    r = np.empty(n_sites, dtype=np.float64)
    lib().orc_xyh_init_uniforms(seed & 0xFFFFFFFF, draw, n_sites, _p(r))
    return r


def xy_init_uniforms(seed: int, draw: int, nx: int, ny: int):
    r = np.empty(nx * ny, dtype=np.float64)
    lib().orc_xy_init_uniforms(seed & 0xFFFFFFFF, draw, nx, ny, _p(r))
    return r


# --------------------------------------------------------------------------
# variance_covariance_kahan (the drivers' accumulator; restated from its use, see oracle.c)
# --------------------------------------------------------------------------
class variance_covariance_kahan:
    def __init__(self):
        self.st = np.zeros(10, dtype=np.float64)
        self.n = 0

    def add_data(self, v1, v2):
        lib().orc_kahan_add_data(_p(self.st), float(v1), float(v2))
        self.n += 1

    def results(self):
        """[num_sample, mean1, mean2, square_mean1, square_mean2, var1, var2, cov]"""
        out = np.empty(8, dtype=np.float64)
        lib().orc_kahan_results(_p(self.st), self.n, _p(out))
        return out


# --------------------------------------------------------------------------
# Ising 2D  (type ising2d_gpu, src/ising2d_gpu_m.f90:12-42)
# --------------------------------------------------------------------------
class ising2d_gpu:
    """CPU restatement of ``type(ising2d_gpu)``; ``rng='philox'`` draws the
    B200 build's counter-based stream, or pass uniforms explicitly to
    ``update(randoms=...)``."""

    def init(self, nx, ny, kbt, iseed):
        self.nx_, self.ny_ = int(nx), int(ny)
        self.nall_ = self.nx_ * self.ny_
        self.seed_ = int(iseed)
        self.draw_ = 0
        self.s = np.empty(self.nall_ + 2 * self.nx_, dtype=np.int32)
        self.set_allup_spin()
        self.set_kbt(kbt)
        return self

    def skip_draws(self, n):
        self.draw_ += int(n)

    def set_allup_spin(self):
        lib().orc_ising2d_set_allup(self.nx_, self.ny_, _p(self.s))

    def _next_uniforms(self, fn=ising_uniforms):
        u = fn(self.seed_, self.draw_, self.nall_)
        self.draw_ += 1
        return u

    def set_random_spin(self, randoms=None):
        if randoms is None:
            randoms = self._next_uniforms(ring_init_uniforms)
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        lib().orc_ising2d_set_random(self.nx_, self.ny_, _p(self.s), _p(randoms))

    def set_kbt(self, kbt):
        self.set_beta(1 / kbt)

    def set_beta(self, beta):
        self.beta_ = float(beta)
        self.exparr = np.empty(17, dtype=np.float64)
        lib().orc_ising2d_exparr(self.beta_, _p(self.exparr))
        self.pup = np.empty(5, dtype=np.float64)
        lib().orc_heatbath_table(self.beta_, 4, _p(self.pup))

    def update(self, randoms=None):
        if randoms is None:
            randoms = self._next_uniforms()
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        lib().orc_ising2d_update(self.nx_, self.ny_, _p(self.s), _p(randoms), _p(self.exparr))

    def update_heatbath(self, randoms=None):
        if randoms is None:
            randoms = self._next_uniforms()
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        lib().orc_ising2d_update_heatbath(self.nx_, self.ny_, _p(self.s), _p(randoms), _p(self.pup))

    def nx(self): return self.nx_
    def ny(self): return self.ny_
    def nall(self): return self.nall_
    def kbt(self): return 1 / self.beta_
    def beta(self): return self.beta_
    def spins(self): return self.s.copy()
    def calc_energy_sum(self): return int(lib().orc_ising2d_energy(self.nx_, self.ny_, _p(self.s)))
    def calc_magne_sum(self): return int(lib().orc_ising2d_magne(self.nx_, self.ny_, _p(self.s)))


# --------------------------------------------------------------------------
# Periodic Ising 2D / 3D (torus): not in the reference -- the "1024^3 periodic" input of BASELINE.md / SURVEY 8(d),
# with the reference's update rule, tables, value conventions (3D 0/1, 2D -1/+1) and observables
# --------------------------------------------------------------------------
def isingp_uniforms(seed, draw, nx, ny, nz, init=False):
    out = np.empty(nx * ny * max(nz, 1), dtype=np.float64)
    (lib().orc_isingp_init_uniforms if init else lib().orc_isingp_uniforms)(seed & 0xFFFFFFFF, draw, nx, ny, max(nz, 1), _p(out))
    return out


class ising_periodic_gpu:
    """CPU restatement of the periodic Ising module (cuda_fortran_mc_simulation_spin_b200/ising_periodic_gpu_m.py);
    nz = 0 -> 2D.  Spins: array [nz][ny][nx] (3D) / [ny][nx] (2D), no halo."""

    def init(self, nx, ny, nz, kbt, iseed):
        self.nx_, self.ny_, self.nz_ = int(nx), int(ny), int(nz)
        self.ndim_ = 3 if self.nz_ > 0 else 2
        self.nall_ = self.nx_ * self.ny_ * max(self.nz_, 1)
        self.seed_ = int(iseed)
        self.draw_ = 0
        self.s = np.empty(self.nall_, dtype=np.int32)
        self.set_allup_spin()
        self.set_kbt(kbt)
        return self

    def skip_draws(self, n):
        self.draw_ += int(n)

    def set_allup_spin(self):
        self.s[:] = 1

    def set_random_spin(self, randoms=None):
        if randoms is None:
            randoms = isingp_uniforms(self.seed_, self.draw_, self.nx_, self.ny_, self.nz_, init=True)
            self.draw_ += 1
        up = np.asarray(randoms) < 0.5                      # src/ising3d_gpu_m.f90:99, src/ising2d_gpu_m.f90:83
        self.s[:] = np.where(up, 1, 0 if self.ndim_ == 3 else -1)

    def set_kbt(self, kbt):
        self.set_beta(1 / kbt)

    def set_beta(self, beta):
        self.beta_ = float(beta)
        if self.ndim_ == 3:
            et = np.empty(8, dtype=np.int64)
            self.w = np.empty(14, dtype=np.float64)
            lib().orc_ising3d_tables(self.beta_, _p(et), _p(self.w))
        else:
            self.w = np.empty(17, dtype=np.float64)
            lib().orc_ising2d_exparr(self.beta_, _p(self.w))
        self.pup = np.empty(7, dtype=np.float64)
        lib().orc_heatbath_table(self.beta_, 6 if self.ndim_ == 3 else 4, _p(self.pup))

    def _next(self, randoms):
        if randoms is None:
            randoms = isingp_uniforms(self.seed_, self.draw_, self.nx_, self.ny_, self.nz_)
            self.draw_ += 1
        return np.ascontiguousarray(randoms, dtype=np.float64)

    def update(self, randoms=None):
        r = self._next(randoms)
        lib().orc_isingp_update(self.ndim_, self.nx_, self.ny_, max(self.nz_, 1), _p(self.s), _p(r), _p(self.w), 0)

    def update_heatbath(self, randoms=None):
        r = self._next(randoms)
        lib().orc_isingp_update(self.ndim_, self.nx_, self.ny_, max(self.nz_, 1), _p(self.s), _p(r), _p(self.pup), 1)

    def nall(self): return self.nall_
    def spins(self): return self.s.copy()

    def measure(self):
        e, m = C.c_int64(0), C.c_int64(0)
        lib().orc_isingp_energy_magne(self.ndim_, self.nx_, self.ny_, max(self.nz_, 1), _p(self.s), C.byref(e), C.byref(m))
        return int(e.value), int(m.value)

    def calc_energy_sum(self): return self.measure()[0]
    def calc_magne_sum(self): return self.measure()[1]


# --------------------------------------------------------------------------
# Ising 3D  (type ising3d_gpu, src/ising3d_gpu_m.f90:15-48)
# --------------------------------------------------------------------------
class ising3d_gpu:
    def init(self, nx, ny, nz, kbt, iseed):
        self.nx_, self.ny_, self.nz_ = int(nx), int(ny), int(nz)
        self.nxy_ = self.nx_ * self.ny_
        self.nall_ = self.nxy_ * self.nz_
        self.seed_ = int(iseed)
        self.draw_ = 0
        self.s = np.empty(self.nall_ + 2 * self.nxy_, dtype=np.int32)
        self.set_allup_spin()
        self.set_kbt(kbt)
        return self

    def skip_draws(self, n):
        self.draw_ += int(n)

    def set_allup_spin(self):
        lib().orc_ising3d_set_allup(self.nx_, self.ny_, self.nz_, _p(self.s))

    def _next_uniforms(self, fn=ising_uniforms):
        u = fn(self.seed_, self.draw_, self.nall_)
        self.draw_ += 1
        return u

    def set_random_spin(self, randoms=None):
        if randoms is None:
            randoms = self._next_uniforms(ring_init_uniforms)
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        lib().orc_ising3d_set_random(self.nx_, self.ny_, self.nz_, _p(self.s), _p(randoms))

    def set_kbt(self, kbt):
        self.set_beta(1 / kbt)

    def set_beta(self, beta):
        self.beta_ = float(beta)
        self.et = np.empty(8, dtype=np.int64)
        self.ws = np.empty(14, dtype=np.float64)
        lib().orc_ising3d_tables(self.beta_, _p(self.et), _p(self.ws))
        self.pup = np.empty(7, dtype=np.float64)
        lib().orc_heatbath_table(self.beta_, 6, _p(self.pup))

    def update(self, randoms=None):
        if randoms is None:
            randoms = self._next_uniforms()
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        lib().orc_ising3d_update(self.nx_, self.ny_, self.nz_, _p(self.s), _p(randoms), _p(self.ws))

    def update_heatbath(self, randoms=None):
        if randoms is None:
            randoms = self._next_uniforms()
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        lib().orc_ising3d_update_heatbath(self.nx_, self.ny_, self.nz_, _p(self.s), _p(randoms), _p(self.pup))

    def nx(self): return self.nx_
    def ny(self): return self.ny_
    def nz(self): return self.nz_
    def nall(self): return self.nall_
    def kbt(self): return 1 / self.beta_
    def beta(self): return self.beta_
    def spins(self): return self.s.copy()

    def calc_energy_sum(self):
        return int(lib().orc_ising3d_energy(self.nx_, self.ny_, self.nz_, _p(self.s), _p(self.et)))

    def calc_magne_sum(self):
        return int(lib().orc_ising3d_magne(self.nx_, self.ny_, self.nz_, _p(self.s)))


# --------------------------------------------------------------------------
# Clock, helical (type clock_gpu, src/clock_gpu_m.f90:13-47; batched twin
# src/clock_gpu_multi_m.f90:13-48 via n_multi / strict comparator)
# --------------------------------------------------------------------------
class clock_gpu:
    def init(self, nx, ny, kbt, state, iseed, n_multi=None):
        self.nx_, self.ny_, self.q_ = int(nx), int(ny), int(state)
        self.nall_ = self.nx_ * self.ny_
        self.multi_ = n_multi is not None
        self.n_multi_ = int(n_multi) if self.multi_ else 1
        self.seed_ = int(iseed)
        self.draw_ = 0
        self.s = np.zeros((self.n_multi_, self.nall_ + 2 * self.nx_), dtype=np.int32)
        self.set_kbt(kbt)
        return self

    def skip_draws(self, n):
        self.draw_ += int(n)

    def set_allup_spin(self):
        self.s[...] = 0

    def set_random_spin(self, randoms=None):
        for j in range(self.n_multi_):
            if randoms is None:
                # replica j uses seed stream (TAG_INIT, draw) with replica folded into the seed
                r = ring_init_uniforms(self.seed_ + 0x9E3779B9 * j, self.draw_, self.nall_)
            else:
                r = np.ascontiguousarray(np.asarray(randoms).reshape(self.n_multi_, -1)[j], dtype=np.float64)
            lib().orc_clock_set_random(self.nx_, self.ny_, self.q_, _p(self.s[j]), _p(r))
        if randoms is None:
            self.draw_ += 1

    def set_kbt(self, kbt):
        self.set_beta(1 / kbt)

    def set_beta(self, beta):
        q = self.q_
        self.beta_ = float(beta)
        self.magne = np.empty(q, dtype=np.float64)
        self.etab = np.empty(q ** 3, dtype=np.float64)
        self.ws = np.empty(q ** 6, dtype=np.float64)
        lib().orc_clock_tables(q, self.beta_, _p(self.magne), _p(self.etab), _p(self.ws))

    def update(self, randoms=None, next_states=None):
        for j in range(self.n_multi_):
            if randoms is None:
                r, p = clock_uniforms(self.seed_, self.draw_, j, self.nall_, self.q_)
            else:
                r = np.ascontiguousarray(np.asarray(randoms).reshape(self.n_multi_, -1)[j], dtype=np.float64)
                p = np.ascontiguousarray(np.asarray(next_states).reshape(self.n_multi_, -1)[j], dtype=np.float64)
            lib().orc_clock_update(self.nx_, self.ny_, self.q_, _p(self.s[j]), _p(r), _p(p),
                                   _p(self.ws), 1 if self.multi_ else 0)
        if randoms is None:
            self.draw_ += 1

    def nx(self): return self.nx_
    def ny(self): return self.ny_
    def nall(self): return self.nall_
    def kbt(self): return 1 / self.beta_
    def beta(self): return self.beta_
    def spins(self): return self.s.copy() if self.multi_ else self.s[0].copy()

    def calc_energy_sum(self):
        r = [lib().orc_clock_energy(self.nx_, self.ny_, self.q_, _p(self.s[j]), _p(self.etab))
             for j in range(self.n_multi_)]
        return np.array(r) if self.multi_ else r[0]

    def calc_magne_sum(self):
        r = [lib().orc_clock_magne(self.nx_, self.ny_, self.q_, _p(self.s[j]), _p(self.magne))
             for j in range(self.n_multi_)]
        return np.array(r) if self.multi_ else r[0]

    def histograms(self, j=0):
        q = self.q_
        hist = np.zeros(q, dtype=np.int64)
        pair = np.zeros(q * q, dtype=np.int64)
        lib().orc_clock_histograms(self.nx_, self.ny_, q, _p(self.s[j]), _p(hist), _p(pair))
        return hist, pair.reshape(q, q)


# --------------------------------------------------------------------------
# Periodic 6-state clock, module-procedure API
# (src/clock/clock_tableall_gpu_m.f90:43-45) as an object
# --------------------------------------------------------------------------
class clock_tableall:
    def __init__(self, nx, ny, kbt, mstate=6):
        self.nx, self.ny, self.q = int(nx), int(ny), int(mstate)
        self.nall = self.nx * self.ny
        self.beta = 1 / kbt
        q = self.q
        self.magne = np.empty(q, dtype=np.float64)
        self.e3 = np.empty(q ** 3, dtype=np.float64)
        self.prob = np.empty(q ** 6, dtype=np.float64)
        lib().orc_tableall_tables(q, self.beta, _p(self.magne), _p(self.e3), _p(self.prob))
        self.c = np.zeros(self.nall, dtype=np.int32)  # sixclock(nx, ny), column-major flat

    def init_sixclock_order(self):
        self.c[...] = 0

    def update_metropolis(self, rnds):
        rnds = np.ascontiguousarray(rnds, dtype=np.float64)
        assert rnds.size == 2 * self.nall
        lib().orc_tableall_update(self.nx, self.ny, self.q, _p(self.c), _p(rnds), _p(self.prob))

    def calc_magne(self):
        return lib().orc_tableall_magne(self.nx, self.ny, self.q, _p(self.c), _p(self.magne))

    def calc_energy(self):
        return lib().orc_tableall_energy(self.nx, self.ny, self.q, _p(self.c), _p(self.e3))

    def histograms(self):
        q = self.q
        hist = np.zeros(q, dtype=np.int64)
        pair = np.zeros(q * q, dtype=np.int64)
        lib().orc_tableall_histograms(self.nx, self.ny, q, _p(self.c), _p(hist), _p(pair))
        return hist, pair.reshape(q, q)


class clock_simple(clock_tableall):
    """src/clock/clock_simple_gpu_m.f90: same model, delta-E summed over the four neighbours on the fly"""

    def __init__(self, nx, ny, kbt, mstate=6):
        super().__init__(nx, ny, kbt, mstate)
        lib().orc_clock_simple_prob(self.q, self.beta, _p(self.prob))


class clock_dual_lattice(clock_tableall):
    """src/clock/clock_dual_lattice_tableall_m.f90: same model, compact colour arrays."""

    def __init__(self, nx, ny, kbt, mstate=6):
        super().__init__(nx, ny, kbt, mstate)
        self.even = np.zeros(self.nall // 2, dtype=np.int32)
        self.odd = np.zeros(self.nall // 2, dtype=np.int32)

    def init_sixclock_order(self):
        self.even[...] = 0
        self.odd[...] = 0

    def update_metropolis(self, rnds):
        rnds = np.ascontiguousarray(rnds, dtype=np.float64)
        lib().orc_dual_update(self.nx, self.ny, self.q, _p(self.even), _p(self.odd), _p(rnds), _p(self.prob))

    def calc_energy(self):
        return lib().orc_dual_energy(self.nx, self.ny, self.q, _p(self.even), _p(self.odd), _p(self.e3))

    def to_full(self):
        """scatter the two colour arrays back to sixclock(nx, ny) (column-major flat)"""
        nx, ny, nh = self.nx, self.ny, self.nx // 2
        full = np.empty((ny, nx), dtype=np.int32)
        ev = self.even.reshape(ny, nh)
        od = self.odd.reshape(ny, nh)
        for y in range(1, ny + 1):
            # even array (parity_bit 0): actual_x = 2X - ((y + 0) & 1)
            X = np.arange(1, nh + 1)
            full[y - 1, 2 * X - (y & 1) - 1] = ev[y - 1]
            full[y - 1, 2 * X - ((y + 1) & 1) - 1] = od[y - 1]
        return full.reshape(-1)


# --------------------------------------------------------------------------
# XY periodic (type xy2d_gpu, src/xy2d_periodic_gpu_m.f90:14-59)
# --------------------------------------------------------------------------
class xy2d_gpu:
    def init(self, nx, ny, kbt, iseed):
        self.nx_, self.ny_ = int(nx), int(ny)
        self.nall_ = self.nx_ * self.ny_
        self.seed_ = int(iseed)
        self.draw_ = 0
        # spins(0:nx+1, 0:ny+1, 1:2) column-major == C array [2][ny+2][nx+2]
        self.sp = np.zeros((2, self.ny_ + 2, self.nx_ + 2), dtype=np.float64)
        self.sp0 = np.zeros_like(self.sp)
        self.set_allup_spin()
        self.set_kbt(kbt)
        return self

    def set_allup_spin(self):
        lib().orc_xy_set_allup(self.nx_, self.ny_, _p(self.sp))

    def set_random_spin(self, randoms, refresh=True):
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        lib().orc_xy_set_random(self.nx_, self.ny_, _p(self.sp), _p(randoms), 1 if refresh else 0)

    def set_angles(self, theta):
        """test helper: interior spins from an (ny, nx) array of angles, halo refreshed"""
        th = np.asarray(theta, dtype=np.float64).reshape(self.ny_, self.nx_)
        self.sp[0, 1:-1, 1:-1] = np.cos(th)
        self.sp[1, 1:-1, 1:-1] = np.sin(th)
        lib().orc_xy_norishiro(self.nx_, self.ny_, _p(self.sp))

    def set_kbt(self, kbt):
        self.set_beta(1 / kbt)

    def set_beta(self, beta):
        self.beta_ = float(beta)

    def set_initial_magne_autocorrelation_state(self):
        self.sp0[...] = self.sp

    def update(self, randoms, candidates):
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        candidates = np.ascontiguousarray(candidates, dtype=np.float64)
        lib().orc_xy_update(self.nx_, self.ny_, _p(self.sp), self.beta_, _p(randoms), _p(candidates))

    def metropolis_by_field(self, randoms, candidates, hx, hy):
        """metropolis_by_field_sub, src/xy2d_periodic_gpu_m.f90:198-216 (no halo refresh, like the reference)"""
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        candidates = np.ascontiguousarray(candidates, dtype=np.float64)
        lib().orc_xy_metropolis_by_field(self.nx_, self.ny_, _p(self.sp), _p(randoms), _p(candidates), float(hx), float(hy))

    def update_over_relaxation(self, n_steps):
        lib().orc_xy_over_relaxation(self.nx_, self.ny_, _p(self.sp), int(n_steps))

    def rotate(self, theta):
        lib().orc_xy_rotate(self.nx_, self.ny_, _p(self.sp), float(theta))
        lib().orc_xy_norishiro(self.nx_, self.ny_, _p(self.sp))

    def nx(self): return self.nx_
    def ny(self): return self.ny_
    def nall(self): return self.nall_
    def kbt(self): return 1 / self.beta_
    def beta(self): return self.beta_
    def spins(self): return self.sp.copy()
    def calc_energy_sum(self): return lib().orc_xy_energy(self.nx_, self.ny_, _p(self.sp))
    def calc_magne_sum(self): return lib().orc_xy_magne(self.nx_, self.ny_, _p(self.sp), 1)
    def calc_magne_y_sum(self): return lib().orc_xy_magne(self.nx_, self.ny_, _p(self.sp), 2)
    def calc_autocorrelation_sum(self): return lib().orc_xy_autocorrelation(self.nx_, self.ny_, _p(self.sp), _p(self.sp0))
    def calc_correlation_sum(self): return lib().orc_xy_correlation(self.nx_, self.ny_, _p(self.sp))


# --------------------------------------------------------------------------
# XY helical (module xy2d_gpu_m, src/xy2d_gpu_m.f90:12-43)
# --------------------------------------------------------------------------
class xy2d_helical_gpu:
    def init(self, nx, ny, kbt, iseed):
        self.nx_, self.ny_ = int(nx), int(ny)
        self.nall_ = self.nx_ * self.ny_
        self.seed_ = int(iseed)
        # spins(1-nx : nall+nx, 1:2) column-major == C array [2][nall + 2 nx]
        self.sp = np.zeros((2, self.nall_ + 2 * self.nx_), dtype=np.float64)
        self.set_allup_spin()
        self.set_kbt(kbt)
        return self

    def set_allup_spin(self): lib().orc_xyh_set_allup(self.nx_, self.ny_, _p(self.sp))

    def set_random_spin(self, randoms):
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        lib().orc_xyh_set_random(self.nx_, self.ny_, _p(self.sp), _p(randoms))

    def set_angles(self, theta):
        """test helper: interior spins from nall angles (linear index order), halo refreshed"""
        th = np.asarray(theta, dtype=np.float64).reshape(-1)
        nx = self.nx_
        self.sp[0, nx:nx + self.nall_] = np.cos(th)
        self.sp[1, nx:nx + self.nall_] = np.sin(th)
        lib().orc_xyh_norishiro(self.nx_, self.ny_, _p(self.sp))

    def set_kbt(self, kbt): self.beta_ = 1 / float(kbt)
    def set_beta(self, beta): self.beta_ = float(beta)

    def update(self, randoms, candidates):
        randoms = np.ascontiguousarray(randoms, dtype=np.float64)
        candidates = np.ascontiguousarray(candidates, dtype=np.float64)
        lib().orc_xyh_update(self.nx_, self.ny_, _p(self.sp), self.beta_, _p(randoms), _p(candidates))

    def update_over_relaxation(self, n_steps):
        lib().orc_xyh_over_relaxation(self.nx_, self.ny_, _p(self.sp), int(n_steps))

    def nx(self): return self.nx_
    def ny(self): return self.ny_
    def nall(self): return self.nall_
    def kbt(self): return 1 / self.beta_
    def beta(self): return self.beta_
    def spins(self): return self.sp.copy()
    def calc_energy_sum(self): return lib().orc_xyh_energy(self.nx_, self.ny_, _p(self.sp))
    def calc_magne_sum(self): return lib().orc_xyh_magne(self.nx_, self.ny_, _p(self.sp))
