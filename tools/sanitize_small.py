"""tiny runs of every model for compute-sanitizer (memcheck / racecheck / initcheck)"""
import sys
sys.path.insert(0, ".")
import numpy as np
from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m, ising3d_gpu_m, clock_gpu_m, clock_gpu_multi_m, xy2d_periodic_gpu_m, xy2d_gpu_m
from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
for shape in [(31, 31, 30), (63, 65, 64), (15, 17, 64)]:
    g = ising3d_gpu_m.ising3d_gpu().init(*shape, 4.51152, 42)
    g.set_random_spin(); g.update_n(3); g.measure(); g.update(); g.measure(); g.run_relaxation(3); s = g.spins(); g.set_spins(s)
    g.set_method(1); g.update(); g.measure()
for shape in [(33, 32), (255, 256), (101, 100)]:
    g = ising2d_gpu_m.ising2d_gpu().init(*shape, 2.269, 42)
    g.set_random_spin(); g.update_n(3); g.measure(); g.update(); g.measure(); g.run_relaxation(3); s = g.spins(); g.set_spins(s)
    g.update_with_randoms(1.0 - np.random.default_rng(1).random(g.nall()))
c = clock_gpu_m.clock_gpu().init(33, 32, 0.9, 6, 42); c.set_random_spin(); c.update_n(2); c.calc_energy_sum(); c.spins()
c = clock_gpu_multi_m.clock_gpu().init(33, 32, 0.9, 6, 2, 42); c.update_n(2); c.calc_energy_sum()
for shape in [(34, 6), (64, 8), (2, 2)]:
    s6 = sixclock(*shape, 0.91, 6, 2, 42); s6.update_metropolis_n(3); s6.calc_energy(); s6.get_sixclock(); s6.get_dual(); s6.close()
x = xy2d_periodic_gpu_m.xy2d_gpu().init(64, 36, 0.89, 42); x.set_random_spin(); x.update_n(2); x.update_over_relaxation(2); x.measure(); x.spins()
x.set_initial_magne_autocorrelation_state(); x.calc_autocorrelation_sum(); x.calc_correlation_sum(); x.metropolis_by_field(1.0, 0.5)
xh = xy2d_gpu_m.xy2d_gpu().init(33, 10, 0.89, 42); xh.set_random_spin(); xh.update_n(2); xh.update_over_relaxation(2); xh.calc_energy_sum(); xh.spins()
# round 2: bit-packed Ising (plain + fused-measurement pass, import / export, halo), class-table clock paths, ragged XY rows
for dim, shape, kbt in [(3, (15, 17, 256), 4.51152), (2, (33, 256), 2.269)]:
    g = (ising3d_gpu_m.ising3d_gpu() if dim == 3 else ising2d_gpu_m.ising2d_gpu()).init_packed(*shape, kbt, 42)
    g.set_random_spin(); g.update_n(2); g.measure(); g.update(); g.measure(); g.update(); g.measure(); s = g.spins(); g.set_spins(s); g.update()
import os
os.environ["B200MC_SIX_DIRECT"] = "0"; os.environ["B200MC_CLOCK_DIRECT"] = "0"
s6 = sixclock(34, 6, 0.91, 6, 2, 42); s6.update_metropolis_n(2); s6.calc_energy(); s6.close()
s6 = sixclock(36, 8, 0.7, 8, 1, 42); s6.update_metropolis_n(2); s6.calc_energy(); s6.close()
c = clock_gpu_m.clock_gpu().init(33, 32, 0.9, 6, 42); c.set_random_spin(); c.update_n(2); c.calc_energy_sum()
c = clock_gpu_m.clock_gpu().init(33, 32, 0.9, 8, 42); c.set_random_spin(); c.update_n(2); c.calc_energy_sum()
x = xy2d_periodic_gpu_m.xy2d_gpu().init(1002, 8, 0.89, 42); x.set_random_spin(); x.update_n(2); x.update_over_relaxation(2); x.measure(); x.update(); x.measure()
print("sanitize run complete")
