"""GPU parity ON THE CODE PATH THE BENCHMARK TIMES.

Lattices of more than (resident blocks x 1024) vectors per colour -- ~14.5 M sites in 3D, ~19.4 M in 2D -- are updated by
the ordered-ticket instantiation `ising_pass_kernel<NNB, METHOD, ORDERED=1>` (and, once the caller measures every MCS,
its fused-measurement twin): the kernel in bench.py's `roofline.kernel` and `e2e`.  The small-lattice tests of
test_gpu_ising.py never reach it (they run the cooperative sweep kernel or the static round-robin pass), so these tests
compare it with the CPU oracle directly: spins (halo cells included) bit-exact, int64 E and M exact, Metropolis and
heat-bath, plain and fused passes; the same under B200MC_TUNE=16, where ONE GPU runs the slab kernels of the
multi-GPU path against its own arrays (`ising_slab_kernel`: boundary tickets + stores into the "neighbour's" halo +
flag handshake, interior blocks on shared tickets); and the headline 1023 x 1023 x 1024 lattice itself.
Reference: update_sub, src/ising3d_gpu_m.f90:174-206; src/ising2d_gpu_m.f90:133-162.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KBT3 = 4.51152
KBT2 = 2.26918531421

# (dim, shape): both have Nc % 16 == 0 (needed by the slab kernels) and are above the ticket threshold
LARGE = [(3, (255, 255, 320)), (2, (4097, 6144))]


def _pair(oracle, dim, shape, seed):
    from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m, ising3d_gpu_m
    if dim == 3:
        return ising3d_gpu_m.ising3d_gpu().init(*shape, KBT3, seed), oracle.ising3d_gpu().init(*shape, KBT3, seed)
    return ising2d_gpu_m.ising2d_gpu().init(*shape, KBT2, seed), oracle.ising2d_gpu().init(*shape, KBT2, seed)


def _launches():
    import ctypes as C
    from cuda_fortran_mc_simulation_spin_b200 import _lib
    return _lib.fn("b200mc_launch_count", C.c_ulonglong)()


@pytest.fixture(params=["plain", "slab_self"])
def kernel_path(request, monkeypatch):
    """plain: ising_pass_kernel<ORDERED> + halo kernel.  slab_self (B200MC_TUNE bit 4): the multi-GPU slab pass with this
    GPU as its own neighbour -- ising_slab_kernel (PUSH boundary tickets, flags) + ising_pass_kernel on shared tickets."""
    if request.param == "slab_self":
        monkeypatch.setenv("B200MC_TUNE", "16")
    return request.param


@pytest.mark.parametrize("dim,shape", LARGE)
@pytest.mark.parametrize("method", [0, 1])
def test_ticket_path_bit_exact(oracle, dim, shape, method, kernel_path):
    g, o = _pair(oracle, dim, shape, 42)
    nvec = g.nall() // 32
    assert nvec > (592 if dim == 2 else 444) * 1024, "shape must be above the ordered-ticket threshold"
    g.set_method(method)
    step = o.update_heatbath if method else o.update
    g.set_random_spin(); o.set_random_spin()
    assert np.array_equal(g.spins(), o.spins())
    # sweep 1: plain pass, E/M by the separate kernel; sweeps 2..4: the second colour pass accumulates E/M itself
    for sweep in range(4):
        l0 = _launches()
        g.update(); step()
        em = g.measure()
        l1 = _launches()
        assert em == (o.calc_energy_sum(), o.calc_magne_sum()), (sweep, em)
        if sweep >= 1 and kernel_path == "plain":
            assert l1 - l0 == 4, "fused sweep = 2 colour passes + 2 halo refreshes, no measure kernel"
        if sweep in (0, 3):
            assert np.array_equal(g.spins(), o.spins()), f"spins differ after sweep {sweep + 1}"
    # update_n: unfused passes followed by one fused pass
    g.update_n(3); step(); step(); step()
    assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())
    assert np.array_equal(g.spins(), o.spins())


@pytest.mark.parametrize("dim,shape", LARGE)
def test_ticket_path_allup_and_series(oracle, dim, shape, kernel_path):
    """all-up start (the drivers' start) + the device-side driver loop (run_relaxation) on the ticket path"""
    g, o = _pair(oracle, dim, shape, 7)
    e, m = g.run_relaxation(3)
    for i in range(3):
        o.update()
        assert (int(e[i]), int(m[i])) == (o.calc_energy_sum(), o.calc_magne_sum()), i
    assert np.array_equal(g.spins(), o.spins())


def test_ticket_path_tail_positions(oracle):
    """a large fold whose last positions hold no site in the high lanes (16 does not divide Nc): masked fused sums"""
    g, o = _pair(oracle, 3, (255, 257, 322), 42)
    assert (g.nall() // 2) % 16 != 0
    g.set_random_spin(); o.set_random_spin()
    for sweep in range(3):
        g.update(); o.update()
        assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum()), sweep
    assert np.array_equal(g.spins(), o.spins())


def test_headline_lattice_against_oracle(oracle):
    """BASELINE config 2 at the shape bench.py runs (1023 x 1023 x 1024, all-up start, kbt = 4.51152, seed 42):
    two sweeps + E/M after each (second one fused), whole configuration bit-exact against the CPU oracle.
    Host memory: ~4.3 GB oracle spins + 8.6 GB uniforms + 4.3 GB export."""
    import os
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 1 << 40
    if avail < 28 * (1 << 30):
        pytest.skip("needs ~28 GB of free host memory")
    g, o = _pair(oracle, 3, (1023, 1023, 1024), 42)
    u = np.empty(o.nall(), dtype=np.float64)
    for sweep in range(2):
        g.update()
        oracle.ising_uniforms_fast(42, sweep, o.nall(), out=u)
        o.update(randoms=u)
        assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum()), sweep
    del u
    s = g.spins()
    assert np.array_equal(s, o.s), "headline lattice: configuration differs from the oracle after 2 sweeps"


def test_set_spins_rejects_invalid_values(oracle):
    """values the byte-parallel kernels cannot hold are refused (B200MC_ERR_ARG), not truncated to a byte"""
    from cuda_fortran_mc_simulation_spin_b200 import B200MCError, clock_gpu_m, ising2d_gpu_m, ising3d_gpu_m
    g3 = ising3d_gpu_m.ising3d_gpu().init(7, 5, 6, KBT3, 1)
    s = g3.spins()
    s[7 * 5 + 3] = 2
    with pytest.raises(B200MCError):
        g3.set_spins(s)
    s[7 * 5 + 3] = -1
    with pytest.raises(B200MCError):
        g3.set_spins(s)
    g2 = ising2d_gpu_m.ising2d_gpu().init(7, 6, KBT2, 1)
    s = g2.spins()
    s[7 + 2] = 0
    with pytest.raises(B200MCError):
        g2.set_spins(s)
    s[7 + 2] = -1
    g2.set_spins(s)
    assert g2.spins()[7 + 2] == -1
    gc = clock_gpu_m.clock_gpu().init(7, 6, 0.8, 6, 1)
    s = gc.spins()
    s[7 + 1] = 6
    with pytest.raises(B200MCError):
        gc.set_spins(s)
    s[7 + 1] = 5
    gc.set_spins(s)
