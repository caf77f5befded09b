"""helical q = 6 clock (clock_gpu_m) at 16385 x 16384: flips/ns of update_n"""
import os, sys
sys.path.insert(0, ".")
import torch
from cuda_fortran_mc_simulation_spin_b200 import clock_gpu_m
c = clock_gpu_m.clock_gpu().init(16385, 16384, 0.91, 6, 42)
c.update_n(3); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); c.update_n(10); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"clock helical DIRECT={os.environ.get('B200MC_CLOCK_DIRECT')}: {ms:.3f} ms/MCS {16385*16384/ms/1e6:.1f} flips/ns  E,M = {c.calc_energy_sum():.3f} {c.calc_magne_sum():.3f}", flush=True)
