"""CPU: the oracle's clock / XY restatements against known answers and against each other
(the reference has no tests, these pin the restatement)."""
import math

import numpy as np


def test_tableall_equals_dual_lattice_and_simple_tables(oracle):
    nx, ny = 34, 6
    a, b, c = oracle.clock_tableall(nx, ny, 0.91, 6), oracle.clock_dual_lattice(nx, ny, 0.91, 6), oracle.clock_simple(nx, ny, 0.91, 6)
    # clock_simple sums delta-E over the four neighbours: same table up to the last bits, not bit-identical
    assert not np.array_equal(a.prob, c.prob) and np.allclose(a.prob, c.prob, rtol=1e-14, atol=0)
    assert a.prob.max() == 1.0 and a.prob.min() > 0.0
    for s in range(5):
        r = oracle.torus_uniforms(42, s, 0, nx, ny)
        assert r.min() > 0.0 and r.max() <= 1.0
        a.update_metropolis(r); b.update_metropolis(r)
        assert np.array_equal(a.c, b.to_full())          # dual lattice indexes its randoms by the full-lattice coordinate
        assert abs(a.calc_energy() - b.calc_energy()) < 1e-12
    h, pair = a.histograms()
    assert h.sum() == nx * ny and pair.sum() == 2 * nx * ny


def test_clock_known_answers(oracle):
    o = oracle.clock_tableall(16, 8, 0.91, 6)
    assert abs(o.calc_energy() + 2.0) < 1e-12 and abs(o.calc_magne() - 1.0) < 1e-12   # ordered state, per site
    cold = oracle.clock_tableall(16, 8, 1e-3, 6)
    cold.update_metropolis(oracle.torus_uniforms(1, 0, 0, 16, 8))
    assert not cold.c.any()                                                            # beta -> infinity: nothing moves
    h = oracle.clock_gpu().init(33, 32, 0.8, 6, 42)
    assert abs(h.calc_energy_sum() + 2.0 * h.nall()) < 1e-9 and abs(h.calc_magne_sum() - h.nall()) < 1e-9


def test_xy_helical_known_answers(oracle):
    o = oracle.xy2d_helical_gpu().init(9, 8, 0.9, 1)
    n = o.nall()
    assert (o.calc_energy_sum(), o.calc_magne_sum()) == (-2.0 * n, 1.0 * n)
    o.set_random_spin(oracle.xyh_init_uniforms(1, 0, n))
    e0 = o.calc_energy_sum()
    o.update_over_relaxation(3)
    assert abs(o.calc_energy_sum() - e0) < 1e-9 * n                 # microcanonical
    # halo rows mirror the interior (update_norishiro_sub, src/xy2d_gpu_m.f90:114-125)
    nx = 9
    assert np.array_equal(o.sp[:, :nx], o.sp[:, n:n + nx]) and np.array_equal(o.sp[:, n + nx:], o.sp[:, nx:2 * nx])
    o.set_beta(1e9)
    before = o.sp.copy()
    r, c = oracle.xyh_uniforms(1, 1, n)
    o.set_allup_spin(); o.update(r, c)
    assert o.calc_energy_sum() == -2.0 * n                          # beta -> infinity from all-up: no candidate accepted
    assert before.shape == o.sp.shape


def test_xy_periodic_known_answers(oracle):
    o = oracle.xy2d_gpu().init(16, 8, 0.89, 3)
    n = o.nall()
    assert (o.calc_energy_sum(), o.calc_magne_sum(), o.calc_magne_y_sum()) == (-2.0 * n, 1.0 * n, 0.0)
    o.set_random_spin(oracle.xy_init_uniforms(3, 0, 16, 8))
    e0 = o.calc_energy_sum()
    o.update_over_relaxation(2)
    assert abs(o.calc_energy_sum() - e0) < 1e-9 * n
    # metropolis_by_field: accepted iff r <= 1 - exp(dE): a field along +x can only raise Mx
    mx0 = o.calc_magne_sum()
    r, c = oracle.xy_uniforms(3, 1, 16, 8)
    o.metropolis_by_field(r, c, 2.0, 0.0)
    assert o.calc_magne_sum() >= mx0


def test_batched_ising_uniforms(oracle):
    """sample 0 of a batch draws the plain stream; other samples differ (sample index = high word of the counter)"""
    a = oracle.ising_uniforms(42, 3, 4096)
    assert np.array_equal(a, oracle.ising_uniforms_rep(42, 3, 0, 4096))
    b = oracle.ising_uniforms_rep(42, 3, 1, 4096)
    assert not np.array_equal(a, b) and abs(b.mean() - 0.5) < 0.03
    assert np.array_equal(oracle.ring_init_uniforms(42, 0, 512), oracle.ring_init_uniforms_rep(42, 0, 0, 512))


def test_kahan_accumulator_against_exact_sums(oracle):
    """the drivers' variance_covariance_kahan, restated from its use: against exact rational arithmetic"""
    from fractions import Fraction
    rng = np.random.default_rng(3)
    acc = oracle.variance_covariance_kahan()
    v1 = 0.5 + 1e-9 * rng.standard_normal(2000)
    v2 = -1.75 + 1e-7 * rng.standard_normal(2000)
    for a, b in zip(v1, v2):
        acc.add_data(a, b)
    r = acc.results()
    n = len(v1)
    f1 = [Fraction(float(a)) for a in v1]
    f2 = [Fraction(float(b)) for b in v2]
    m1, m2 = sum(f1) / n, sum(f2) / n
    assert r[0] == n
    assert abs(r[1] - float(m1)) <= 2e-16 and abs(r[2] - float(m2)) <= 4e-16
    assert abs(r[3] - float(sum(a * a for a in f1) / n)) <= 2e-16
    cov = (sum(a * b for a, b in zip(f1, f2)) / n - m1 * m2) * Fraction(n, n - 1)
    var1 = (sum(a * a for a in f1) / n - m1 * m1) * Fraction(n, n - 1)
    # the variance of a nearly constant series is a difference of nearly equal means: absolute accuracy ~ 1 ulp of the square mean
    assert abs(r[5] - float(var1)) <= 4e-16 and abs(r[7] - float(cov)) <= 1e-15
    one = oracle.variance_covariance_kahan()
    one.add_data(0.25, 0.5)
    assert one.results().tolist() == [1.0, 0.25, 0.5, 0.0625, 0.25, 0.0, 0.0, 0.0]
