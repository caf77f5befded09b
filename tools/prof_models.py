"""a few sweeps of every model at the BASELINE sizes (ncu target: tools/prof_models.py <model>)"""
import sys
sys.path.insert(0, ".")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("ising3d", "all"):
    from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m
    m = ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42)
    m.update_n(3); print("ising3d", m.measure())
    for _ in range(3):
        m.update(); m.measure()          # fused-measurement variant of the second colour pass
    del m
if which in ("ising2d", "all"):
    from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m
    m = ising2d_gpu_m.ising2d_gpu().init(65537, 65536, 2.26918531421, 42)
    m.update_n(3); print("ising2d", m.measure()); del m
if which in ("xy", "all"):
    from cuda_fortran_mc_simulation_spin_b200 import xy2d_periodic_gpu_m as xm
    g = xm.xy2d_gpu().init(16384, 16384, 0.89, 42); g.set_random_spin()
    g.update_n(3); g.update_over_relaxation(3); print("xy", g.measure()); del g
if which in ("sixclock", "all"):
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
    g = sixclock(16384, 16384, 0.91, 6, 2, 42)
    g.update_metropolis_n(3); print("sixclock", g.calc_energy(), g.calc_magne()); g.close()
if which in ("clock", "all"):
    from cuda_fortran_mc_simulation_spin_b200 import clock_gpu_m
    c = clock_gpu_m.clock_gpu().init(16385, 16384, 0.91, 6, 42)
    c.update_n(3); print("clock", c.calc_energy_sum(), c.calc_magne_sum()); del c
if which in ("ising_small",):
    from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m
    m = ising2d_gpu_m.ising2d_gpu().init(1001, 1000, 2.26918531421, 42)
    m.update_n(3); m.update_n(100); print("ising_small", m.measure()); del m
if which in ("ising_slabself",):   # run with B200MC_TUNE=16: the slab pass against the GPU's own arrays
    from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m
    m = ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42)
    m.update_n(4); print("ising_slabself", m.measure()); del m
