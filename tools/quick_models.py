import sys, os
sys.path.insert(0, ".")
import torch
from cuda_fortran_mc_simulation_spin_b200 import xy2d_periodic_gpu_m as xm, clock_gpu_m, clock_gpu_multi_m
def ev(fn, n):
    fn(3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
g = xm.xy2d_gpu().init(16384, 16384, 0.89, 42); g.set_random_spin()
ms = ev(g.update_n, 20); print(f"xy metropolis 16384^2: {ms:.3f} ms  {g.nall()/ms/1e6:.1f} flips/ns ({g.nall()/ms/1e6*12/6456.5*100:.1f}% of 12 B/flip roofline)")
ms = ev(g.update_over_relaxation, 20); print(f"xy over-relax  16384^2: {ms:.3f} ms  {g.nall()/ms/1e6:.1f} flips/ns")
def meas(n):
    for _ in range(n): g.update(); g.measure()
ms = ev(meas, 10); print(f"xy update+measure: {ms:.3f} ms")
del g
c = clock_gpu_m.clock_gpu().init(16385, 16384, 0.91, 6, 42)
ms = ev(c.update_n, 10); print(f"clock q=6 16385x16384: {ms:.3f} ms  {c.nall()/ms/1e6:.1f} flips/ns")
def measc(n):
    for _ in range(n): c.update(); c.calc_energy_sum(); c.calc_magne_sum()
ms = ev(measc, 5); print(f"clock update+measure: {ms:.3f} ms")
del c
c = clock_gpu_multi_m.clock_gpu().init(16385, 16384, 0.91, 6, 2, 42)
ms = ev(c.update_n, 5); print(f"clock multi n=2: {ms:.3f} ms  {2*c.nall()/ms/1e6:.1f} flips/ns")
del c
c = clock_gpu_m.clock_gpu().init(1001, 1000, 0.8, 6, 42)
ms = ev(c.update_n, 200); print(f"clock q=6 1001x1000 (reference's published config): {ms*1e3:.1f} us  {c.nall()/ms/1e6:.1f} flips/ns  (reference note: ~10 flips/ns)")
