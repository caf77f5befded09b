"""GPU parity on the REFERENCE'S OWN RANDOM STREAM.

The reference fills its uniform arrays with cuRAND: `curandCreateGenerator(gen, CURAND_RNG_PSEUDO_XORWOW)`,
`curandSetPseudoRandomGeneratorSeed(gen, iseed)` at init and `curandGenerate(gen, real64_device_array, n)` per sweep
(src/ising3d_gpu_m.f90:64-65,179; src/ising2d_gpu_m.f90:56-57,138; src/clock_gpu_m.f90:73-74,188-189;
src/clock/clock_tableall_gpu_m.f90:64-65,95; src/xy2d_periodic_gpu_m.f90:74-75,355-356).  The reference itself cannot
be built here (CUDA Fortran), but its generator can be driven: these tests create that generator through libcurand's
host API on the GPU box, generate the arrays with the call sequence of the cited lines, and feed the SAME arrays to the
CPU oracle and to the CUDA path's reference-stream entry points (`update_with_randoms` / `update_with_rnds`).  This pins
the kernels and the oracle to each other on the reference's stream (value range (0, 1], ties at table entries, index
order of the arrays) -- it is the one pin this image allows; it does not replace running the reference.
"""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CURAND_RNG_PSEUDO_XORWOW = 101


class Xorwow:
    """the reference's generator: XORWOW, host API, default ordering, double-precision uniforms"""

    def __init__(self, seed):
        import torch
        self.torch = torch
        lib = None
        for name in ("libcurand.so.10", "/usr/local/cuda/lib64/libcurand.so.10", "libcurand.so"):
            try:
                lib = C.CDLL(name)
                break
            except OSError:
                continue
        if lib is None:
            pytest.skip("libcurand not found")
        self.lib = lib
        self.gen = C.c_void_p(None)
        assert lib.curandCreateGenerator(C.byref(self.gen), C.c_int(CURAND_RNG_PSEUDO_XORWOW)) == 0
        assert lib.curandSetPseudoRandomGeneratorSeed(self.gen, C.c_ulonglong(seed)) == 0

    def set_offset(self, n):
        """curandSetGeneratorOffset (skip_curand, src/ising3d_gpu_m.f90:72-77)"""
        assert self.lib.curandSetGeneratorOffset(self.gen, C.c_ulonglong(n)) == 0

    def generate(self, n):
        """curandGenerate(gen, real64 array, n) of the nvfortran curand module = curandGenerateUniformDouble"""
        t = self.torch.empty(int(n), dtype=self.torch.float64, device="cuda")
        assert self.lib.curandGenerateUniformDouble(self.gen, C.c_void_p(t.data_ptr()), C.c_size_t(int(n))) == 0
        self.torch.cuda.synchronize()
        return t.cpu().numpy()

    def __del__(self):
        try:
            if self.gen:
                self.lib.curandDestroyGenerator(self.gen)
        except Exception:
            pass


def test_curand_stream_properties():
    """what the kernels may assume about the reference's uniforms: 0 < u <= 1 (the author's own comment,
    src/clock/clock_tableall_gpu_m.f90:140), reproducible per seed, and skip_curand's offset selects a later part of
    the same stream"""
    n = 1 << 20
    a = Xorwow(42).generate(n)
    assert a.min() > 0.0 and a.max() <= 1.0
    assert abs(a.mean() - 0.5) < 2e-3
    assert np.array_equal(a, Xorwow(42).generate(n))
    assert not np.array_equal(a, Xorwow(43).generate(n))
    two = Xorwow(42)
    first, second = two.generate(n), two.generate(n)
    assert np.array_equal(first, a)            # successive generate calls continue the stream
    assert not np.array_equal(second, a)
    g = Xorwow(42)
    g.set_offset(n)                             # skip_curand: a different part of the stream
    b = g.generate(n)
    assert b.min() > 0.0 and b.max() <= 1.0 and not np.array_equal(a, b)


@pytest.mark.parametrize("shape", [(31, 31, 30), (63, 65, 64)])
def test_ising3d_on_curand_stream(oracle, shape):
    from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m
    g = ising3d_gpu_m.ising3d_gpu().init(*shape, 4.51152, 42)
    o = oracle.ising3d_gpu().init(*shape, 4.51152, 42)
    gen = Xorwow(42)
    n = g.nall()
    o.set_random_spin(gen.generate(n))       # set_random_spin draws from the same generator, :86
    g.set_spins(o.spins())
    for sweep in range(5):
        u = gen.generate(n)                   # :179
        g.update_with_randoms(u)
        o.update(randoms=u)
        assert np.array_equal(g.spins(), o.spins()), sweep
        assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())


def test_ising2d_on_curand_stream(oracle):
    from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m
    g = ising2d_gpu_m.ising2d_gpu().init(1001, 1000, 2.26918531421, 42)   # app/ising2d_gpu_relaxation.f90 defaults
    o = oracle.ising2d_gpu().init(1001, 1000, 2.26918531421, 42)
    gen = Xorwow(42)
    for sweep in range(5):
        u = gen.generate(g.nall())            # src/ising2d_gpu_m.f90:138
        g.update_with_randoms(u)
        o.update(randoms=u)
        assert np.array_equal(g.spins(), o.spins()), sweep
        assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())


@pytest.mark.parametrize("q", [6, 4])
def test_clock_on_curand_stream(oracle, q):
    from cuda_fortran_mc_simulation_spin_b200 import clock_gpu_m
    g = clock_gpu_m.clock_gpu().init(501, 500, 0.8, q, 42)
    o = oracle.clock_gpu().init(501, 500, 0.8, q, 42)
    gen = Xorwow(42)
    n = g.nall()
    for sweep in range(4):
        r = gen.generate(n)                   # src/clock_gpu_m.f90:188
        p = gen.generate(n)                   # :189
        g.update_with_randoms(r, p)
        o.update(r, p)
        assert np.array_equal(g.spins(), o.spins()), sweep
        assert abs(g.calc_energy_sum() - o.calc_energy_sum()) <= 1e-11 * n
        assert abs(g.calc_magne_sum() - o.calc_magne_sum()) <= 1e-12 * n


def test_tableall_on_curand_stream(oracle):
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
    nx, ny = 256, 128
    g = sixclock(nx, ny, 0.91, 6, 1, 42)
    o = oracle.clock_tableall(nx, ny, 0.91, 6)
    gen = Xorwow(42)
    for sweep in range(4):
        rnds = gen.generate(2 * nx * ny)      # rnds(2, nx, ny), src/clock/clock_tableall_gpu_m.f90:95
        g.update_with_rnds(rnds)
        o.update_metropolis(rnds)
        assert np.array_equal(g.get_sixclock()[0], o.c), sweep
        assert abs(g.calc_energy()[0] - o.calc_energy()) <= 1e-12
        assert abs(g.calc_magne()[0] - o.calc_magne()) <= 1e-12


def test_xy_on_curand_stream(oracle):
    """XY: fp32 angles and SFU math against the real64 oracle, per sweep from the same state, on the reference's stream
    (randoms first, then candidates: src/xy2d_periodic_gpu_m.f90:355-356).  Tolerance as in test_gpu_xy.py: 1e-5
    relative on E, Mx, My (floor sqrt(N): the scale of a disordered lattice's sums) plus the counted borderline
    accept decisions."""
    from cuda_fortran_mc_simulation_spin_b200 import xy2d_periodic_gpu_m as xm
    from test_gpu_xy import check_observables, sync_oracle
    nx, ny = 256, 128
    g = xm.xy2d_gpu().init(nx, ny, 0.89, 42)
    o = oracle.xy2d_gpu().init(nx, ny, 0.89, 42)
    gen = Xorwow(42)
    n = nx * ny
    o.set_random_spin(gen.generate(n))        # :107
    g.set_angles(np.arctan2(o.sp[1, 1:-1, 1:-1], o.sp[0, 1:-1, 1:-1]) / (2 * math.pi))
    for sweep in range(4):
        sync_oracle(o, g)
        r, c = gen.generate(n), gen.generate(n)
        g.update_with_randoms(r, c)
        o.update(r, c)
        check_observables(g, o, n)
