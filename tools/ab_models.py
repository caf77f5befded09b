import os, sys
sys.path.insert(0, ".")
import torch
from cuda_fortran_mc_simulation_spin_b200 import xy2d_periodic_gpu_m as xm
from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
def ev(fn, n):
    fn(3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
g = xm.xy2d_gpu().init(16384, 16384, 0.89, 42); g.set_random_spin()
a = ev(g.update_n, 20); b = ev(g.update_over_relaxation, 20); n = g.nall(); del g
s = sixclock(16384, 16384, 0.91, 6, 2, 42)
c = ev(s.update_metropolis_n, 10)
print(f"{os.path.basename(os.environ.get('B200MC_SO','default'))}: xy metropolis {n/a/1e6:.1f}  over-relax {n/b/1e6:.1f}  sixclock x2 {2*n/c/1e6:.1f} flips/ns", flush=True)
