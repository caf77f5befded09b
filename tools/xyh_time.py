"""helical XY (xy2d_gpu_m) at 16385 x 16384: flips/ns of update / update_over_relaxation"""
import sys
sys.path.insert(0, ".")
import torch
from cuda_fortran_mc_simulation_spin_b200 import xy2d_gpu_m
g = xy2d_gpu_m.xy2d_gpu().init(16385, 16384, 0.89, 42); g.set_random_spin()
def ev(fn, n):
    fn(2); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def upd(n):
    for _ in range(n): g.update()
a = ev(upd, 8); b = ev(g.update_over_relaxation, 8)
print(f"xy helical 16385x16384: metropolis {g.nall()/a/1e6:.1f}  over-relax {g.nall()/b/1e6:.1f} flips/ns", flush=True)
