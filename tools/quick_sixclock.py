"""time the periodic 6-state clock (tableall / dual lattice) at BASELINE config 4: 16384^2, batch of samples"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
for (nx, ny, nm) in [(16384, 16384, 1), (16384, 16384, 4), (2000, 2000, 1), (2000, 2000, 16)]:
    g = sixclock(nx, ny, 0.91, 6, nm, 42)
    g.update_metropolis_n(3); g.sync()
    n = 10 if nx > 4000 else 50
    g.set_timing(True)
    t0 = time.perf_counter(); g.update_metropolis_n(n); g.sync(); dt = time.perf_counter() - t0
    nl, ms = g.get_timing()
    g.set_timing(False)
    t1 = time.perf_counter(); e = g.calc_energy(); m = g.calc_magne(); dm = time.perf_counter() - t1
    print(f"sixclock {nx}x{ny} x{nm}: {dt/n*1e3:.3f} ms/MCS  {nx*ny*nm*n/dt/1e9:.1f} flips/ns  (kernel {ms/nl:.3f} ms/launch, "
          f"{nx*ny*nm/2/(ms/nl)/1e6:.1f} flips/ns in-kernel)  measure {dm*1e3:.2f} ms  e={e[0]:.6f} m={m[0]:.6f}", flush=True)
    g.close()
