import sys, os
sys.path.insert(0, ".")
import torch
from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m, ising2d_gpu_m
def timeit(m, n):
    m.update_n(3); m.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    m.update_n(n)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
m = ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42)
ms = timeit(m, 30)
print(f"TUNE={os.environ.get('B200MC_TUNE')} ising3d: {ms:.3f} ms/sweep  {m.nall()/ms/1e6:.1f} flips/ns  E,M={m.measure()}", flush=True)
del m
m = ising2d_gpu_m.ising2d_gpu().init(65537, 65536, 2.26918531421, 42)
ms = timeit(m, 10)
print(f"  ising2d 65537x65536: {ms:.3f} ms/sweep  {m.nall()/ms/1e6:.1f} flips/ns", flush=True)
