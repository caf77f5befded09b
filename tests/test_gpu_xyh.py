"""GPU parity: XY with the helical boundary (module xy2d_gpu_m) vs the real64 CPU oracle,
per sweep from a shared state, 1e-5 relative (fp32 angles on the GPU)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _sync(o, g):
    o.set_angles(2 * math.pi * g.angles().astype(np.float64))


@pytest.mark.parametrize("shape,kbt", [((5, 4), 0.89), ((33, 32), 0.895), ((255, 64), 0.5), ((1001, 500), 0.89), ((7, 6), 1.5)])
def test_xyh_metropolis_and_over_relaxation(oracle, shape, kbt):
    from cuda_fortran_mc_simulation_spin_b200 import xy2d_gpu_m as xm
    nx, ny = shape
    n = nx * ny
    g = xm.xy2d_gpu().init(nx, ny, kbt, 42)
    o = oracle.xy2d_helical_gpu().init(nx, ny, kbt, 42)
    assert (g.calc_energy_sum(), g.calc_magne_sum()) == (-2.0 * n, 1.0 * n)       # all-up known answer
    g.set_random_spin()
    assert np.allclose(g.angles(), oracle.xyh_init_uniforms(42, 0, n), atol=2 ** -24)
    for sweep in range(4):
        _sync(o, g)
        assert abs(g.calc_energy_sum() - o.calc_energy_sum()) <= RTOL * n and abs(g.calc_magne_sum() - o.calc_magne_sum()) <= RTOL * n
        r, c = oracle.xyh_uniforms(42, 1 + sweep, n)
        g.update(); o.update(r, c)
        assert abs(g.calc_energy_sum() - o.calc_energy_sum()) <= RTOL * max(abs(o.calc_energy_sum()), 0.05 * n) + (2e-3 * n if n < 100 else 0)
        assert abs(g.calc_magne_sum() - o.calc_magne_sum()) <= RTOL * n + (2e-3 * n if n < 100 else 0)
        go = g.angles().astype(np.float64)
        oo = np.arctan2(o.sp[1, nx:nx + n], o.sp[0, nx:nx + n]) / (2 * math.pi)
        d = np.abs(((go - oo + 0.5) % 1.0) - 0.5)
        assert (d > 1e-5).mean() < max(1e-4, 1.5 / n)
    # over-relaxation: microcanonical, and equal to the oracle's reflection
    g.update_n(2)
    for it in range(2):
        _sync(o, g)
        e0 = g.calc_energy_sum()
        g.update_over_relaxation(1); o.update_over_relaxation(1)
        assert abs(g.calc_energy_sum() - e0) <= 1e-5 * n
        assert abs(g.calc_energy_sum() - o.calc_energy_sum()) <= RTOL * max(abs(o.calc_energy_sum()), 0.05 * n)


def test_xyh_spins_layout_and_limits(oracle):
    from cuda_fortran_mc_simulation_spin_b200 import B200MCError, xy2d_gpu_m as xm
    nx, ny = 33, 16
    g = xm.xy2d_gpu().init(nx, ny, 0.89, 3)
    o = oracle.xy2d_helical_gpu().init(nx, ny, 0.89, 3)
    g.set_random_spin()
    _sync(o, g)
    assert np.allclose(g.spins(), o.sp, atol=1e-12)          # interior and both halo rows
    a = g.angles(); g.set_angles(a); assert np.array_equal(g.angles(), a)
    g.set_allup_spin(); g.set_beta(1e9); g.update_n(2)
    assert g.calc_energy_sum() == -2.0 * nx * ny             # beta -> inf from all-up: nothing accepted
    with pytest.raises(B200MCError):
        xm.xy2d_gpu().init(32, 16, 0.89, 3)                  # even nx: linear-index colouring is not a checkerboard
