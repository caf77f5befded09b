"""CPU: the oracle's Ising 2D/3D restatement against known answers
and against independent brute-force computations."""
import math

import numpy as np
import pytest

KBT2 = 2.26918531421
KBT3 = 4.51152


def test_tables_match_closed_forms(oracle):
    o = oracle.ising3d_gpu().init(5, 5, 4, KBT3, 1)
    beta = 1 / KBT3
    for S in range(7):
        assert o.ws[S + 7 * 0] == min(1.0, math.exp(-beta * (12 - 4 * S)))   # src/ising3d_gpu_m.f90:163
        assert o.ws[S + 7 * 1] == min(1.0, math.exp(-beta * (4 * S - 12)))   # :164
    assert o.et.tolist() == [-3, -1, 1, 3, 3, 1, -1, -3]                      # energy_table(S3, s) = -sigma (2 S3 - 3)
    o2 = oracle.ising2d_gpu().init(5, 4, KBT2, 1)
    assert all(o2.exparr[d + 8] == 1.0 for d in range(-8, 1))
    assert all(o2.exparr[d + 8] == math.exp(-(1 / KBT2) * d) for d in range(1, 9))


def _brute_energy_3d(s, nx, ny, nz):
    """-sum over bonds of sigma sigma' on the helical ring, computed from the interior only"""
    n = nx * ny * nz
    sig = 2 * s.astype(np.int64) - 1
    e = 0
    for d in (1, nx, nx * ny):
        e -= int((sig * np.roll(sig, -d)).sum())
    return e


def test_known_answers_allup(oracle):
    o = oracle.ising3d_gpu().init(7, 5, 6, KBT3, 1)
    n = o.nall()
    assert (o.calc_energy_sum(), o.calc_magne_sum()) == (-3 * n, n)
    o2 = oracle.ising2d_gpu().init(7, 6, KBT2, 1)
    assert (o2.calc_energy_sum(), o2.calc_magne_sum()) == (-2 * o2.nall(), o2.nall())


def test_beta_limits(oracle):
    o = oracle.ising3d_gpu().init(9, 7, 8, KBT3, 3)
    n = o.nall()
    o.set_beta(1e6)
    for _ in range(3):
        o.update()
    assert (o.calc_energy_sum(), o.calc_magne_sum()) == (-3 * n, n)      # nothing flips
    o.set_beta(0.0)
    o.update()
    assert o.calc_magne_sum() == -n                                       # everything flips
    o2 = oracle.ising2d_gpu().init(9, 8, KBT2, 3)
    o2.set_beta(0.0)
    o2.update()
    assert o2.calc_magne_sum() == -o2.nall()


@pytest.mark.parametrize("shape", [(3, 3, 2), (7, 5, 6), (15, 13, 12)])
def test_energy_vs_bruteforce_3d(oracle, shape):
    nx, ny, nz = shape
    o = oracle.ising3d_gpu().init(nx, ny, nz, KBT3, 11)
    o.set_random_spin()
    for _ in range(3):
        o.update()
        interior = o.s[nx * ny:-nx * ny]
        assert o.calc_energy_sum() == _brute_energy_3d(interior, nx, ny, nz)
        assert o.calc_magne_sum() == int(2 * interior.sum() - o.nall())
        # halo cells are exact copies of the ring continuation
        assert np.array_equal(o.s[:nx * ny], interior[-nx * ny:])
        assert np.array_equal(o.s[-nx * ny:], interior[:nx * ny])


def test_energy_vs_bruteforce_2d(oracle):
    nx, ny = 9, 8
    o = oracle.ising2d_gpu().init(nx, ny, KBT2, 5)
    o.set_random_spin()
    for _ in range(3):
        o.update()
        sig = o.s[nx:-nx].astype(np.int64)
        e = -int((sig * np.roll(sig, -1)).sum()) - int((sig * np.roll(sig, -nx)).sum())
        assert o.calc_energy_sum() == e and o.calc_magne_sum() == int(sig.sum())


def test_colour_pass_is_order_independent(oracle):
    """within one colour only the other colour is read (valid shapes): a serial
    pass in reversed site order gives the same result as the oracle's forward/OpenMP pass"""
    nx, ny, nz = 7, 5, 6
    o = oracle.ising3d_gpu().init(nx, ny, nz, KBT3, 2)
    o.set_random_spin()
    s0 = o.spins()
    u = oracle.ising_uniforms(2, 99, o.nall())
    o.update(randoms=u)
    nxy, nall = nx * ny, nx * ny * nz
    s = s0.copy()
    S = lambda idx: s[idx + nxy - 1]
    for offset in (1, 2):
        for idx in range(nall - (nall - offset) % 2, 0, -2):     # reversed order
            tot = S(idx - 1) + S(idx + 1) + S(idx - nx) + S(idx + nx) + S(idx - nxy) + S(idx + nxy)
            if not u[idx - 1] > o.ws[tot + 7 * S(idx)]:
                s[idx + nxy - 1] = 1 - S(idx)
        s[nall + nxy:] = s[nxy:2 * nxy]
        s[:nxy] = s[nall:nall + nxy]
    assert np.array_equal(s, o.spins())


def test_2d_equilibrium_energy_at_tc(oracle):
    """Onsager: e(Tc) = -sqrt(2) in the thermodynamic limit; 129 x 128 after 1500 MCS from all-up
    sits within a few percent (finite size + critical slowing down)"""
    o = oracle.ising2d_gpu().init(129, 128, KBT2, 42)
    for _ in range(1500):
        o.update()
    es = []
    for _ in range(300):
        o.update()
        es.append(o.calc_energy_sum() / o.nall())
    assert abs(np.mean(es) + math.sqrt(2)) < 0.05, np.mean(es)


def test_heatbath_detailed_balance_small(oracle):
    """heat-bath (no reference symbol): p_up / (1 - p_up) = exp(2 beta h)"""
    o = oracle.ising3d_gpu().init(5, 5, 4, KBT3, 1)
    beta = 1 / KBT3
    for S in range(7):
        h = 2 * S - 6
        assert math.isclose(o.pup[S] / (1 - o.pup[S]), math.exp(2 * beta * h), rel_tol=1e-12)
    # long heat-bath and Metropolis runs agree on <e> at a high temperature (fast mixing)
    a = oracle.ising2d_gpu().init(33, 32, 3.5, 7)
    b = oracle.ising2d_gpu().init(33, 32, 3.5, 8)
    ea, eb = [], []
    for i in range(600):
        a.update(); b.update_heatbath()
        if i >= 100:
            ea.append(a.calc_energy_sum() / a.nall()); eb.append(b.calc_energy_sum() / b.nall())
    assert abs(np.mean(ea) - np.mean(eb)) < 0.02
