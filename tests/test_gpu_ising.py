"""GPU parity: Ising 2D / 3D CUDA path (through the C ABI) vs the CPU oracle.

Bar (BASELINE.json north_star): spin configurations and the int64 energy /
magnetisation sums are BIT-EXACT.  The oracle consumes the same uniforms the
kernels generate in registers (oracle/rng_contract.c) or an explicit array
(update_with_randoms), in the reference's own layouts (halo cells included).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KBT3 = 4.51152          # app/ising3d_gpu_relaxation.f90:12
KBT2 = 2.26918531421    # app/ising2d_gpu_relaxation.f90:11


def _mods():
    from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m, ising3d_gpu_m
    return ising2d_gpu_m, ising3d_gpu_m


def test_device_philox_kat():
    """Random123 known-answer vectors for Philox4x32-10, on the device."""
    import ctypes as C
    from cuda_fortran_mc_simulation_spin_b200 import _lib
    f = _lib.fn("b200mc_debug_philox", C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        _lib.check(f(c.ctypes.data, k.ctypes.data, o.ctypes.data))
        assert o.tolist() == want


@pytest.fixture(params=["coop", "launches"])
def pass_path(request, monkeypatch):
    """small lattices run whole sweeps in one cooperative launch (ising_coop_kernel); B200MC_TUNE bit 11 forces the
    launch-per-colour-pass path that large lattices use, so that both are checked against the oracle"""
    if request.param == "launches":
        monkeypatch.setenv("B200MC_TUNE", "2048")
    return request.param


SHAPES3 = [(3, 3, 2), (5, 5, 4), (7, 5, 6), (31, 31, 30), (33, 31, 34), (63, 65, 64), (101, 101, 100)]


@pytest.mark.parametrize("shape", SHAPES3)
@pytest.mark.parametrize("start", ["allup", "random"])
def test_ising3d_trajectory_bit_exact(oracle, shape, start, pass_path):
    _, i3 = _mods()
    nx, ny, nz = shape
    g = i3.ising3d_gpu().init(nx, ny, nz, KBT3, 42)
    o = oracle.ising3d_gpu().init(nx, ny, nz, KBT3, 42)
    assert np.array_equal(g.ws(), o.ws.reshape(2, 7))
    if start == "random":
        g.set_random_spin()
        o.set_random_spin()
    assert np.array_equal(g.spins(), o.spins())
    for sweep in range(6):
        g.update()
        o.update()
        assert np.array_equal(g.spins(), o.spins()), f"spins differ after sweep {sweep + 1}"
        assert g.calc_energy_sum() == o.calc_energy_sum()
        assert g.calc_magne_sum() == o.calc_magne_sum()
        assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())


@pytest.mark.parametrize("shape", [(5, 4), (7, 6), (33, 32), (255, 256), (1001, 1000)])
@pytest.mark.parametrize("start", ["allup", "random"])
def test_ising2d_trajectory_bit_exact(oracle, shape, start, pass_path):
    i2, _ = _mods()
    nx, ny = shape
    g = i2.ising2d_gpu().init(nx, ny, KBT2, 42)
    o = oracle.ising2d_gpu().init(nx, ny, KBT2, 42)
    assert np.array_equal(g.exparr(), o.exparr)
    if start == "random":
        g.set_random_spin()
        o.set_random_spin()
    assert np.array_equal(g.spins(), o.spins())
    for sweep in range(6):
        g.update()
        o.update()
        assert np.array_equal(g.spins(), o.spins()), f"spins differ after sweep {sweep + 1}"
        assert g.calc_energy_sum() == o.calc_energy_sum()
        assert g.calc_magne_sum() == o.calc_magne_sum()


def test_ising3d_update_with_randoms_edge_uniforms(oracle):
    """explicit uniform arrays, including u == 1.0, u == table entries and tiny u"""
    _, i3 = _mods()
    nx, ny, nz = 15, 13, 12
    g = i3.ising3d_gpu().init(nx, ny, nz, KBT3, 1)
    o = oracle.ising3d_gpu().init(nx, ny, nz, KBT3, 1)
    rng = np.random.default_rng(7)
    g.set_random_spin(); o.set_random_spin()
    for it in range(4):
        u = 1.0 - rng.random(g.nall())           # (0, 1]
        u[::17] = 1.0
        u[1::19] = o.ws[0]                        # exactly on a table entry: accepted (<=)
        u[2::23] = np.nextafter(o.ws[1], 2.0)     # just above: rejected
        u[3::29] = 1e-300
        g.update_with_randoms(u)
        o.update(randoms=u)
        assert np.array_equal(g.spins(), o.spins())
        assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())


def test_ising2d_update_with_randoms(oracle):
    i2, _ = _mods()
    g = i2.ising2d_gpu().init(101, 100, KBT2, 3)
    o = oracle.ising2d_gpu().init(101, 100, KBT2, 3)
    rng = np.random.default_rng(11)
    for it in range(4):
        u = 1.0 - rng.random(g.nall())
        u[::13] = 1.0
        g.update_with_randoms(u)
        o.update(randoms=u)
        assert np.array_equal(g.spins(), o.spins())
        assert g.calc_energy_sum() == o.calc_energy_sum()


def test_spins_roundtrip_and_observables(oracle):
    i2, i3 = _mods()
    rng = np.random.default_rng(5)
    g = i3.ising3d_gpu().init(21, 19, 18, KBT3, 9)
    o = oracle.ising3d_gpu().init(21, 19, 18, KBT3, 9)
    s = o.spins()
    nxy = 21 * 19
    s[nxy:-nxy] = rng.integers(0, 2, size=o.nall())
    o.s[...] = s
    oracle.lib().orc_ising3d_norishiro(21, 19, 18, o.s.ctypes.data)
    g.set_spins(o.spins())
    assert np.array_equal(g.spins(), o.spins())          # halo cells rebuilt identically
    assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())
    g2 = i2.ising2d_gpu().init(51, 50, KBT2, 9)
    o2 = oracle.ising2d_gpu().init(51, 50, KBT2, 9)
    o2.s[51:-51] = rng.integers(0, 2, size=o2.nall()) * 2 - 1
    oracle.lib().orc_ising2d_norishiro(51, 50, o2.s.ctypes.data)
    g2.set_spins(o2.spins())
    assert np.array_equal(g2.spins(), o2.spins())
    assert g2.measure() == (o2.calc_energy_sum(), o2.calc_magne_sum())


@pytest.mark.parametrize("dim", [2, 3])
def test_heatbath_bit_exact(oracle, dim, pass_path):
    """heat-bath has no reference symbol: kernel vs our own oracle definition"""
    i2, i3 = _mods()
    if dim == 3:
        g = i3.ising3d_gpu().init(31, 31, 30, KBT3, 42); o = oracle.ising3d_gpu().init(31, 31, 30, KBT3, 42)
    else:
        g = i2.ising2d_gpu().init(101, 100, KBT2, 42); o = oracle.ising2d_gpu().init(101, 100, KBT2, 42)
    g.set_method(1)
    for sweep in range(5):
        g.update(); o.update_heatbath()
        assert np.array_equal(g.spins(), o.spins())
        assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())


def test_known_answers_small():
    i2, i3 = _mods()
    g = i3.ising3d_gpu().init(31, 31, 30, KBT3, 42)
    n = g.nall()
    assert g.measure() == (-3 * n, n)                    # all-up: E = -3N, M = N
    g.set_beta(1e6)                                      # beta -> inf from all-up: nothing flips
    g.update_n(3)
    assert g.measure() == (-3 * n, n)
    g.set_beta(0.0)                                      # beta = 0: every proposal accepted
    g.update()
    assert g.measure() == (-3 * n, -n)
    g.update()
    assert g.measure() == (-3 * n, n)
    g2 = i2.ising2d_gpu().init(1001, 1000, KBT2, 42)
    n2 = g2.nall()
    assert g2.measure() == (-2 * n2, n2)
    g2.set_beta(0.0); g2.update()
    assert g2.measure() == (-2 * n2, -n2)


def test_invalid_shapes_rejected():
    from cuda_fortran_mc_simulation_spin_b200 import B200MCError
    i2, i3 = _mods()
    with pytest.raises(B200MCError):
        i3.ising3d_gpu().init(1001, 1000, 1000, KBT3, 42)   # the app default races in the reference (Q1)
    with pytest.raises(B200MCError):
        i3.ising3d_gpu().init(32, 31, 30, KBT3, 42)
    with pytest.raises(B200MCError):
        i2.ising2d_gpu().init(1000, 1000, KBT2, 42)
    with pytest.raises(B200MCError):
        i2.ising2d_gpu().init(1001, 1001, KBT2, 42)


def test_skip_curand_disjoint_streams(oracle):
    _, i3 = _mods()
    g = i3.ising3d_gpu().init(31, 31, 30, KBT3, 42)
    o = oracle.ising3d_gpu().init(31, 31, 30, KBT3, 42)
    g.skip_curand(5 * g.nall()); o.skip_draws(5)
    g.update(); o.update()
    assert np.array_equal(g.spins(), o.spins())


def test_full_size_properties_headline():
    """BASELINE config 2 at the reference-valid shape next to 1024^3: size-independent checks"""
    _, i3 = _mods()
    g = i3.ising3d_gpu().init(1023, 1023, 1024, KBT3, 42)
    n = g.nall()
    assert n == 1023 * 1023 * 1024
    assert g.measure() == (-3 * n, n)
    g.set_beta(0.0); g.update()
    assert g.measure() == (-3 * n, -n)                    # every spin flipped exactly once
    g.set_kbt(KBT3); g.set_allup_spin()
    g.update_n(2)
    e, m = g.measure()
    assert -3 * n < e < -n and 0.5 * n < m < n            # two sweeps from all-up at Tc
    assert (e + 3 * n) % 4 == 0                            # E = -3N + 2X, and X (anti-aligned bonds) is even: every flip toggles 6 bonds


@pytest.mark.parametrize("dim,method", [(3, 0), (3, 1), (2, 0), (2, 1)])
def test_fused_measurement_equals_separate_pass(oracle, dim, method, pass_path):
    """Once the caller measures after an update, the second colour pass of the following sweeps
    accumulates X and sum(s) itself (deferred tie accepts included).  Shapes without site-less tail
    positions (Nc % 16 == 0) take that path; the sums must equal the oracle's and a recount of the
    same configuration by the separate kernel."""
    i2, i3 = _mods()
    if dim == 3:
        g = i3.ising3d_gpu().init(63, 65, 64, KBT3, 5); o = oracle.ising3d_gpu().init(63, 65, 64, KBT3, 5)
    else:
        g = i2.ising2d_gpu().init(255, 256, KBT2, 5); o = oracle.ising2d_gpu().init(255, 256, KBT2, 5)
    assert (g.nall() // 2) % 16 == 0
    g.set_method(method)
    step = o.update_heatbath if method else o.update
    g.set_random_spin(); o.set_random_spin()
    for sweep in range(8):
        if sweep == 5:
            g.update_n(3); step(); step(); step()      # only the last of n sweeps is fused
        else:
            g.update(); step()
        em = g.measure()
        assert em == (o.calc_energy_sum(), o.calc_magne_sum()), sweep
        g.set_spins(g.spins())                         # invalidates: next measure recounts with the separate kernel
        assert g.measure() == em
    g.update(); g.update(); step(); step()             # first result never read: no stale sums afterwards
    assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())


@pytest.mark.parametrize("dim,shape", [(3, (63, 65, 64)), (3, (31, 31, 30)), (2, (255, 256)), (2, (1001, 1000))])
def test_run_relaxation_series(oracle, dim, shape, pass_path):
    """the drivers' loop on the device: per-MCS E and M series == update + measure step by step (fused second-pass
    sums where Nc % 16 == 0, the measure kernel otherwise), and the state afterwards is the same"""
    i2, i3 = _mods()
    if dim == 3:
        g = i3.ising3d_gpu().init(*shape, KBT3, 9); o = oracle.ising3d_gpu().init(*shape, KBT3, 9)
    else:
        g = i2.ising2d_gpu().init(*shape, KBT2, 9); o = oracle.ising2d_gpu().init(*shape, KBT2, 9)
    e, m = g.run_relaxation(6)
    for i in range(6):
        o.update()
        assert (int(e[i]), int(m[i])) == (o.calc_energy_sum(), o.calc_magne_sum()), i
    assert np.array_equal(g.spins(), o.spins())
    assert g.measure() == (int(e[-1]), int(m[-1]))
    g.update(); o.update()
    assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())


def test_tma_staged_pass_bit_exact(oracle, monkeypatch):
    """the copy-engine staged colour pass (ising_pass_tma_kernel, opt-in with B200MC_TUNE bit 7: cp.async.bulk into a
    shared-memory ring, producer warp + 8 consumer warps) on a lattice large enough for the ticket path"""
    i2, i3 = _mods()
    monkeypatch.setenv("B200MC_TUNE", "128")
    shape = (255, 257, 322)
    g = i3.ising3d_gpu().init(*shape, KBT3, 42)
    o = oracle.ising3d_gpu().init(*shape, KBT3, 42)
    g.set_random_spin(); o.set_random_spin()
    for sweep in range(2):
        g.update(); o.update()
        assert np.array_equal(g.spins(), o.spins()), sweep
        assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())


@pytest.mark.parametrize("dim,shape", [(3, (31, 31, 30)), (3, (63, 65, 64)), (2, (255, 256)), (2, (101, 100))])
def test_batch_of_samples_bit_exact(oracle, dim, shape):
    """n_multi independent samples updated by the same launches: sample j == a CPU oracle fed the uniforms of
    sample j (counter high word = j); sample 0 == the plain handle; per-sample E/M and the run_relaxation series"""
    i2, i3 = _mods()
    n = 3
    mod, kbt = (i3.ising3d_gpu, KBT3) if dim == 3 else (i2.ising2d_gpu, KBT2)
    omod = oracle.ising3d_gpu if dim == 3 else oracle.ising2d_gpu
    g = mod().init_multi(*shape, kbt, 42, n)
    assert g.n_multi() == n
    os_ = [omod().init(*shape, kbt, 42) for _ in range(n)]
    nall = g.nall()
    g.set_random_spin()
    for j, o in enumerate(os_):
        o.set_random_spin(oracle.ring_init_uniforms_rep(42, 0, j, nall))
        assert np.array_equal(g.spins_multi(j), o.spins())
    for sweep in range(4):
        g.update()
        e, m = g.measure_multi()
        for j, o in enumerate(os_):
            o.update(randoms=oracle.ising_uniforms_rep(42, 1 + sweep, j, nall))
            assert np.array_equal(g.spins_multi(j), o.spins()), (j, sweep)
            assert (int(e[j]), int(m[j])) == (o.calc_energy_sum(), o.calc_magne_sum()), (j, sweep)
    es, ms = g.run_relaxation(3)
    assert es.shape == (n, 3)
    for i in range(3):
        for j, o in enumerate(os_):
            o.update(randoms=oracle.ising_uniforms_rep(42, 5 + i, j, nall))
            assert (int(es[j, i]), int(ms[j, i])) == (o.calc_energy_sum(), o.calc_magne_sum()), (i, j)
    assert not np.array_equal(g.spins_multi(0), g.spins_multi(1))
    # sample 0 of a batch is the plain handle's trajectory
    p = mod().init(*shape, kbt, 42)
    p.set_random_spin(); p.update_n(7)
    assert np.array_equal(p.spins(), g.spins_multi(0))
    s = os_[2].spins(); g.set_spins_multi(1, s)
    assert np.array_equal(g.spins_multi(1), s)


def test_full_size_properties_large_lattices():
    """size-independent checks at BASELINE config 5 (Ising 2D 65537 x 65536 per GPU) and at an 8.6e9-site 3D lattice
    (2047 x 2047 x 2048: 8 x the headline size, 2 x 4.3 GB of int8 spins)"""
    i2, i3 = _mods()
    g = i2.ising2d_gpu().init(65537, 65536, KBT2, 42)
    n = g.nall()
    assert n == 65537 * 65536 and g.measure() == (-2 * n, n)
    g.set_beta(0.0); g.update()
    assert g.measure() == (-2 * n, -n)
    g.set_kbt(KBT2); g.set_allup_spin(); g.update_n(2)
    e, m = g.measure()
    assert -2 * n < e < -n and 0.5 * n < m < n and (e + 2 * n) % 4 == 0
    del g
    g = i3.ising3d_gpu().init(2047, 2047, 2048, KBT3, 42)
    n = g.nall()
    assert n == 2047 * 2047 * 2048 and g.measure() == (-3 * n, n)
    g.set_beta(0.0); g.update()
    assert g.measure() == (-3 * n, -n)
    g.set_kbt(KBT3); g.set_random_spin()
    e0, m0 = g.measure()
    assert abs(m0) < 1e-3 * n and abs(e0) < 1e-3 * n            # iid Bernoulli(1/2) start
    g.update_n(2)
    e, m = g.measure()
    assert e < e0 and (e + 3 * n) % 4 == 0


@pytest.mark.parametrize("dim,shape,n_multi,random_start", [(3, (31, 31, 30), 1, False), (2, (101, 100), 1, True), (3, (15, 17, 16), 4, False),
                                                             (2, (255, 256), 2, True)])
def test_run_relaxation_stats_equals_the_drivers_loop(oracle, dim, shape, n_multi, random_start):
    """the drivers' whole measurement (app/ising3d_gpu_relaxation.f90:37-55) on the device -- tot_sample x
    [initial state; mcs x (update; m; e; add_data(m / N, e / N))] with Kahan mean / variance / covariance per MCS -- against
    the same loop driven on the CPU oracle with the oracle's restatement of variance_covariance_kahan: same operations in
    the same order, so the eight columns agree to 1e-12 (they are bit-identical unless the compiler contracts differently)."""
    i2, i3 = _mods()
    mcs, tot = 5, 8
    mod, kbt = (i3.ising3d_gpu, KBT3) if dim == 3 else (i2.ising2d_gpu, KBT2)
    omod = oracle.ising3d_gpu if dim == 3 else oracle.ising2d_gpu
    g = mod().init_multi(*shape, kbt, 42, n_multi) if n_multi > 1 else mod().init(*shape, kbt, 42)
    nall = g.nall()
    stats = g.run_relaxation_stats(mcs, tot, random_start)
    acc = [oracle.variance_covariance_kahan() for _ in range(mcs)]
    n_inv = 1.0 / float(nall)
    os_ = [omod().init(*shape, kbt, 42) for _ in range(n_multi)]
    draw = 0
    for batch in range(tot // n_multi):
        if random_start:
            for j, o in enumerate(os_):
                o.set_random_spin(oracle.ring_init_uniforms_rep(42, draw, j, nall))
            draw += 1
        else:
            for o in os_:
                o.set_allup_spin()
        for i in range(mcs):
            for j, o in enumerate(os_):
                o.update(randoms=oracle.ising_uniforms_rep(42, draw, j, nall))
                acc[i].add_data(o.calc_magne_sum() * n_inv, o.calc_energy_sum() * n_inv)
            draw += 1
    want = np.array([a.results() for a in acc])
    assert stats.shape == (mcs, 8) and np.all(stats[:, 0] == tot)
    assert np.allclose(stats, want, rtol=1e-12, atol=1e-15), np.abs(stats - want).max()
    rows = g.format_relaxation_table(stats)
    f = rows[2].split()
    assert len(f) == 10 and int(f[0]) == nall and int(f[1]) == tot and int(f[2]) == 3
    assert float(f[3]) == stats[2, 1] and float(f[7]) == nall * stats[2, 5]
    # the handle goes on from where the loop left it
    g.update()
    for j, o in enumerate(os_):
        o.update(randoms=oracle.ising_uniforms_rep(42, draw, j, nall))
    assert np.array_equal(g.spins_multi(0) if n_multi > 1 else g.spins(), os_[0].spins())
