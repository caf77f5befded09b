"""GPU parity of the periodic (torus) Ising module against its CPU restatement (oracle/oracle.c orc_isingp_*): spins
bit-exact after every sweep, int64 E and M exact (separate kernel and the sums fused into the second colour pass),
all-up and random starts, Metropolis and heat-bath, built-in RNG and explicit uniform arrays (incl. u == 1.0 and
u == a table entry), for both kernels: the generic one (any nx % 32 == 0) and the strip kernel with its rolling row
window (nx % 1024 == 0; 8 and 2 rows per ticket; strips narrower than the row).  Update rule and observables:
src/ising3d_gpu_m.f90:189-206,239-276, src/ising2d_gpu_m.f90:148-162,198-228."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KBT3, KBT2 = 4.51152, 2.26918531421

# (nx, ny, nz); nz = 0 -> 2D.  nx = 1024: one strip per row; 2048: two strips (edge bytes from memory); ny % 8 != 0: 2-row tickets
SHAPES = [(32, 2, 2), (64, 6, 4), (96, 10, 6), (32, 4, 0), (160, 18, 0),
          (1024, 8, 4), (1024, 6, 2), (2048, 16, 2), (1024, 24, 0), (2048, 10, 0), (3072, 8, 0),
          (16384, 8, 0), (65536, 8, 0)]   # the 2D kernels with the row pitch compiled in (R = 512, 2048)


def _pair(oracle, shape, seed=42):
    from cuda_fortran_mc_simulation_spin_b200 import ising_periodic_gpu_m as M
    kbt = KBT3 if shape[2] else KBT2
    return M.ising_periodic_gpu().init(*shape, kbt, seed), oracle.ising_periodic_gpu().init(*shape, kbt, seed)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("start", ["allup", "random"])
def test_torus_trajectory_bit_exact(oracle, shape, method, start):
    g, o = _pair(oracle, shape)
    g.set_method(method)
    step = o.update_heatbath if method else o.update
    assert g.measure() == o.measure()            # all-up: E = -(z / 2) N, M = N
    if start == "random":
        g.set_random_spin(); o.set_random_spin()
    assert np.array_equal(g.spins(), o.spins())
    for sweep in range(5):
        g.update(); step()
        assert np.array_equal(g.spins(), o.spins()), f"spins differ after sweep {sweep + 1}"
        assert g.calc_energy_sum() == o.calc_energy_sum()
        assert g.calc_magne_sum() == o.calc_magne_sum()
    g.update_n(3); step(); step(); step()
    assert g.measure() == o.measure()
    assert np.array_equal(g.spins(), o.spins())


@pytest.mark.parametrize("shape", [(64, 6, 4), (1024, 8, 2), (96, 10, 0), (2048, 8, 0)])
def test_generic_and_strip_kernels_agree(oracle, shape, monkeypatch):
    """the same lattice through the generic kernel (B200MC_TORUS_GENERIC=1) and the default path"""
    g, o = _pair(oracle, shape, 7)
    monkeypatch.setenv("B200MC_TORUS_GENERIC", "1")
    g2, _ = _pair(oracle, shape, 7)
    for m in (g, g2):
        m.set_random_spin()
    o.set_random_spin()
    for _ in range(4):
        g.update(); g2.update(); o.update()
        assert g.measure() == g2.measure() == o.measure()
    assert np.array_equal(g.spins(), g2.spins())
    assert np.array_equal(g.spins(), o.spins())


@pytest.mark.parametrize("shape", [(64, 6, 4), (1024, 8, 2), (96, 10, 0)])
@pytest.mark.parametrize("method", [0, 1])
def test_torus_update_with_randoms_edges(oracle, shape, method):
    """explicit uniforms, compared as real64 like the reference: random values, u == 1.0, u == table entries"""
    g, o = _pair(oracle, shape)
    g.set_method(method)
    g.set_random_spin(); o.set_random_spin()
    rng = np.random.default_rng(5)
    tab = g.table()
    vals = np.unique(np.concatenate([tab.ravel(), [1.0]]))
    vals = vals[(vals > 0) & (vals <= 1.0)]
    for sweep in range(3):
        u = 1.0 - rng.random(g.nall())            # (0, 1]
        pick = rng.random(g.nall()) < 0.3
        u[pick] = rng.choice(vals, size=int(pick.sum()))
        g.update_with_randoms(u)
        (o.update_heatbath if method else o.update)(u)
        assert np.array_equal(g.spins(), o.spins()), f"sweep {sweep}"
        assert g.measure() == o.measure()


def test_torus_table_is_the_reference_table(oracle):
    g, o = _pair(oracle, (64, 6, 4))
    assert np.array_equal(g.table()[:, :7], o.w.reshape(2, 7))          # ws(0:6, 0:1), src/ising3d_gpu_m.f90:153-171
    g2, o2 = _pair(oracle, (64, 6, 0))
    t = g2.table()
    for s in (0, 1):
        for S in range(5):
            assert t[s, S] == o2.w[2 * (2 * s - 1) * (2 * S - 4) + 8]   # exparr(2 s sum), src/ising2d_gpu_m.f90:126-130,195


def test_torus_set_spins_roundtrip_and_validation(oracle):
    from cuda_fortran_mc_simulation_spin_b200 import B200MCError
    g, o = _pair(oracle, (64, 6, 4))
    rng = np.random.default_rng(1)
    s = rng.integers(0, 2, g.nall()).astype(np.int32)
    g.set_spins(s); o.s[:] = s
    assert np.array_equal(g.spins(), s)
    assert g.measure() == o.measure()
    s[5] = 2
    with pytest.raises(B200MCError):
        g.set_spins(s)
    g2, o2 = _pair(oracle, (64, 6, 0))
    s2 = (2 * rng.integers(0, 2, g2.nall()) - 1).astype(np.int32)
    g2.set_spins(s2); o2.s[:] = s2
    assert np.array_equal(g2.spins(), s2)
    assert g2.measure() == o2.measure()


def test_torus_rejects_invalid_shapes():
    from cuda_fortran_mc_simulation_spin_b200 import B200MCError
    from cuda_fortran_mc_simulation_spin_b200 import ising_periodic_gpu_m as M
    for shape in [(48, 4, 4), (64, 5, 4), (64, 4, 3), (0, 4, 4)]:
        with pytest.raises(B200MCError):
            M.ising_periodic_gpu().init(*shape, KBT3, 1)


def test_torus_skip_curand_advances_the_stream(oracle):
    g, o = _pair(oracle, (64, 6, 4))
    g.skip_curand(3 * g.nall() - 5); o.skip_draws(3)
    g.set_random_spin(); o.set_random_spin()
    g.update(); o.update()
    assert np.array_equal(g.spins(), o.spins())


def test_torus_full_size_against_oracle(oracle):
    """the north-star lattice itself, 1024^3: two sweeps + fused E / M against the oracle (hash of the spins)"""
    import hashlib
    g, o = _pair(oracle, (1024, 1024, 1024))
    g.set_random_spin(); o.set_random_spin()
    for _ in range(2):
        g.update(); o.update()
        assert g.measure() == o.measure()
    assert hashlib.sha256(g.spins().tobytes()).hexdigest() == hashlib.sha256(o.spins().tobytes()).hexdigest()
