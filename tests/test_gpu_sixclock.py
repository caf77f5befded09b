"""GPU parity: periodic q-state clock with the q^6 table (clock_tableall_gpu_m and its dual-lattice
twin) vs the CPU oracle, through the C ABI.  States bit-exact after every sweep (tableall array and
dual-lattice colour arrays), integer histograms exact, per-site real64 E and M within 1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_hist(o):
    q, nx, ny = o.q, o.nx, o.ny
    c = o.c.reshape(ny, nx)
    hist = np.bincount(c.ravel(), minlength=q)
    br = np.bincount(((np.roll(c, -1, axis=1) - c) % q).ravel(), minlength=q)
    bu = np.bincount(((np.roll(c, -1, axis=0) - c) % q).ravel(), minlength=q)
    return hist, br, bu


SHAPES = [((2, 2), 6, 0.91), ((4, 4), 6, 0.91), ((32, 32), 6, 0.91), ((34, 6), 6, 0.8), ((30, 8), 6, 0.91),
          ((64, 64), 6, 0.91), ((200, 100), 6, 0.91), ((66, 10), 3, 0.7), ((96, 12), 5, 0.9), ((40, 40), 8, 0.6),
          ((512, 512), 6, 0.91)]


@pytest.fixture(params=["direct", "classes"])
def lookup(request, monkeypatch):
    """acceptance lookup: the one-load u16 threshold table in shared memory (default for q <= 6) or the class-id table +
    per-class thresholds (B200MC_SIX_DIRECT=0; what larger q use) -- both are checked against the oracle"""
    if request.param == "classes":
        monkeypatch.setenv("B200MC_SIX_DIRECT", "0")
    return request.param


@pytest.mark.parametrize("shape,q,kbt", SHAPES)
def test_tableall_trajectory_bit_exact(oracle, shape, q, kbt, lookup):
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
    nx, ny = shape
    g = sixclock(nx, ny, kbt, q, 1, 42)
    o = oracle.clock_tableall(nx, ny, kbt, q)
    assert np.array_equal(g.states_to_prob(), o.prob)     # q^6 real64 table, bit for bit
    assert np.array_equal(g.energy_table(), o.e3)
    assert np.array_equal(g.get_sixclock()[0], o.c)
    assert abs(g.calc_energy()[0] + 2.0) < 1e-12 and abs(g.calc_magne()[0] - 1.0) < 1e-12   # ordered start
    for sweep in range(4):
        g.update_metropolis()
        o.update_metropolis(oracle.torus_uniforms(42, sweep, 0, nx, ny, q))
        assert np.array_equal(g.get_sixclock()[0], o.c), f"states differ after sweep {sweep + 1}"
        h, br, bu = g.histograms()
        oh, obr, obu = _oracle_hist(o)
        assert np.array_equal(h[0], oh) and np.array_equal(br[0], obr) and np.array_equal(bu[0], obu)
        assert abs(g.calc_energy()[0] - o.calc_energy()) <= 1e-12
        assert abs(g.calc_magne()[0] - o.calc_magne()) <= 1e-12


def test_dual_lattice_arrays_bit_exact(oracle):
    """the dual-lattice module's even/odd arrays (src/clock/clock_dual_lattice_tableall_m.f90:22)"""
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
    nx, ny = 68, 20
    g = sixclock(nx, ny, 0.91, 6, 1, 7)
    o = oracle.clock_dual_lattice(nx, ny, 0.91, 6)
    for sweep in range(4):
        g.update_metropolis()
        o.update_metropolis(oracle.torus_uniforms(7, sweep, 0, nx, ny))
        ev, od = g.get_dual()
        assert np.array_equal(ev[0], o.even) and np.array_equal(od[0], o.odd)
        assert np.array_equal(g.get_sixclock()[0], o.to_full())
        assert abs(g.calc_energy()[0] - o.calc_energy()) <= 1e-12
    # set_dual / set_sixclock round trips
    rng = np.random.default_rng(3)
    ev = rng.integers(0, 6, size=nx * ny // 2, dtype=np.int32)
    od = rng.integers(0, 6, size=nx * ny // 2, dtype=np.int32)
    g.set_dual(ev, od)
    e2, o2 = g.get_dual()
    assert np.array_equal(e2[0], ev) and np.array_equal(o2[0], od)
    full = rng.integers(0, 6, size=nx * ny, dtype=np.int32)
    g.set_sixclock(full)
    assert np.array_equal(g.get_sixclock()[0], full)


@pytest.mark.parametrize("shape", [(36, 10), (64, 32), (2, 4)])
def test_update_with_rnds_reference_stream(oracle, shape):
    """arbitrary real64 uniforms in the reference's rnds(2, nx, ny) order, incl. u == 1 and u == a table entry"""
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
    nx, ny = shape
    g = sixclock(nx, ny, 0.91, 6, 1, 1)
    o = oracle.clock_tableall(nx, ny, 0.91, 6)
    rng = np.random.default_rng(11)
    start = rng.integers(0, 6, size=nx * ny, dtype=np.int32)
    g.set_sixclock(start)
    o.c[...] = start
    for sweep in range(3):
        r = 1.0 - rng.random(2 * nx * ny)          # (0, 1]
        r[::7] = 1.0
        r[1::2][::5] = rng.choice(o.prob, size=r[1::2][::5].size)
        g.update_with_rnds(r)
        o.update_metropolis(r)
        assert np.array_equal(g.get_sixclock()[0], o.c)


def test_multi_sample_batch(oracle, lookup):
    """n_multi independent samples in one launch per colour; sample j is the high word of the Philox block counter"""
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
    nx, ny, n = 48, 16, 3
    g = sixclock(nx, ny, 0.8, 6, n, 42)
    os_ = [oracle.clock_tableall(nx, ny, 0.8, 6) for _ in range(n)]
    for sweep in range(4):
        g.update_metropolis()
        s = g.get_sixclock()
        e, m = g.calc_energy(), g.calc_magne()
        for j, o in enumerate(os_):
            o.update_metropolis(oracle.torus_uniforms(42, sweep, j, nx, ny))
            assert np.array_equal(s[j], o.c)
            assert abs(e[j] - o.calc_energy()) <= 1e-12 and abs(m[j] - o.calc_magne()) <= 1e-12
    assert not np.array_equal(s[0], s[1])
    g.init_sixclock_order()
    assert not g.get_sixclock().any()


def test_known_answers_and_skip():
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
    g = sixclock(64, 64, 1e-3, 6, 1, 42)            # beta -> infinity from the ordered state: nothing moves
    g.update_metropolis_n(3)
    assert not g.get_sixclock().any()
    a = sixclock(64, 64, 0.91, 6, 1, 42)
    b = sixclock(64, 64, 0.91, 6, 1, 42)
    a.update_metropolis_n(2)
    b.skip_curand_clock(2 * 64 * 64)                # one update's worth of uniforms (:95)
    b.update_metropolis()
    a2 = sixclock(64, 64, 0.91, 6, 1, 42)
    a2.skip_curand_clock(1)                         # any positive skip advances to the next draw
    a2.update_metropolis()
    assert np.array_equal(a2.get_sixclock(), b.get_sixclock())
    with pytest.raises(Exception):
        sixclock(63, 64, 0.91, 6, 1, 42)            # odd nx: (x + y) colouring is not a checkerboard on the torus


def test_module_procedure_mirrors(oracle):
    """the module-level API of clock_tableall_gpu_m / clock_dual_lattice_tableall_gpu_m (:43-45),
    driven like app/clock_tableall_gpu_relaxation.f90:18-44"""
    from cuda_fortran_mc_simulation_spin_b200 import clock_dual_lattice_tableall_gpu_m as D
    from cuda_fortran_mc_simulation_spin_b200 import clock_tableall_gpu_m as T
    for mod in (T, D):
        mod.configure(nx_=40, ny_=24, kbt_=0.91)
        mod.init_sixclock(42)
        mod.skip_curand_clock(0)
        mod.init_sixclock_order()
    o = oracle.clock_tableall(40, 24, 0.91, 6)
    for i in range(3):
        T.update_metropolis(); D.update_metropolis()
        o.update_metropolis(oracle.torus_uniforms(42, i, 0, 40, 24))
        assert abs(T.calc_magne() - o.calc_magne()) <= 1e-12 and abs(T.calc_energy() - o.calc_energy()) <= 1e-12
        assert T.calc_magne() == D.calc_magne() and T.calc_energy() == D.calc_energy()
    assert np.array_equal(T.sixclock(), o.c)


def test_clock_table_and_simple_variants(oracle):
    """f4: clock_table_gpu_m evaluates tableall's delta-E expression per site (same decisions);
    clock_simple_gpu_m sums over the four neighbours (src/clock/clock_simple_gpu_m.f90:108-113) -- a table
    that differs from tableall's in the last bits, reproduced bit for bit"""
    from cuda_fortran_mc_simulation_spin_b200 import clock_simple_gpu_m as S
    from cuda_fortran_mc_simulation_spin_b200 import clock_table_gpu_m as T
    nx, ny = 72, 20
    for mod in (S, T):
        mod.configure(nx_=nx, ny_=ny, kbt_=0.91)
        mod.init_sixclock(42)
        mod.init_sixclock_order()
    os_, ot = oracle.clock_simple(nx, ny, 0.91, 6), oracle.clock_tableall(nx, ny, 0.91, 6)
    assert np.array_equal(S.handle().states_to_prob(), os_.prob)
    assert np.array_equal(T.handle().states_to_prob(), ot.prob)
    assert not np.array_equal(os_.prob, ot.prob) and np.allclose(os_.prob, ot.prob, rtol=1e-14, atol=0)
    for i in range(4):
        r = oracle.torus_uniforms(42, i, 0, nx, ny)
        S.update_metropolis(); T.update_metropolis()
        os_.update_metropolis(r); ot.update_metropolis(r)
        assert np.array_equal(S.sixclock(), os_.c) and np.array_equal(T.sixclock(), ot.c)
        assert abs(S.calc_energy() - os_.calc_energy()) <= 1e-12 and abs(S.calc_magne() - os_.calc_magne()) <= 1e-12


def test_full_size_properties_config4():
    """BASELINE config 4: q = 6 clock, 16384 x 16384, batch of samples (size-independent checks)"""
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
    g = sixclock(16384, 16384, 0.91, 6, 2, 42)
    n = g.nall()
    assert np.allclose(g.calc_energy(), -2.0) and np.allclose(g.calc_magne(), 1.0)
    g.update_metropolis_n(2)
    h, br, bu = g.histograms()
    assert (h.sum(axis=1) == n).all() and (br.sum(axis=1) == n).all() and (bu.sum(axis=1) == n).all()
    e, m = g.calc_energy(), g.calc_magne()
    assert (-2.0 < e).all() and (e < -1.0).all() and (0.5 < m).all() and (m < 1.0).all()
    assert abs(e[0] - e[1]) < 1e-3 and e[0] != e[1]            # independent samples of the same ensemble
    g.set_kbt(1e-3); g.init_sixclock_order(); g.update_metropolis()
    assert np.allclose(g.calc_energy(), -2.0)                  # beta -> infinity from order: nothing moves


def test_c4_lattice_against_oracle(oracle):
    """BASELINE config 4 at full size (q = 6, 16384 x 16384, kbt = 0.91, two samples per launch): two sweeps of the batch,
    states of both samples bit-exact against the CPU oracle (~1 GB of int32 states + 4.3 GB of uniforms per sample)"""
    import os
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 1 << 40
    if avail < 16 * (1 << 30):
        pytest.skip("needs ~16 GB of free host memory")
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
    nx = ny = 16384
    g = sixclock(nx, ny, 0.91, 6, 2, 42)
    os_ = [oracle.clock_tableall(nx, ny, 0.91, 6) for _ in range(2)]
    for sweep in range(2):
        g.update_metropolis()
        for j, o in enumerate(os_):
            o.update_metropolis(oracle.torus_uniforms(42, sweep, j, nx, ny, 6))
    s = g.get_sixclock()
    e, m = g.calc_energy(), g.calc_magne()
    for j, o in enumerate(os_):
        assert np.array_equal(s[j], o.c), f"sample {j}: states differ from the oracle after 2 sweeps"
        assert abs(e[j] - o.calc_energy()) <= 1e-12 and abs(m[j] - o.calc_magne()) <= 1e-12
    g.close()
