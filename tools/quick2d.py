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
tag = f"TUNE={os.environ.get('B200MC_TUNE')} CHUNK={os.environ.get('B200MC_CHUNK')}"
for shape in [(65537, 65536), (16385, 16384)]:
    m = ising2d_gpu_m.ising2d_gpu().init(*shape, 2.26918531421, 42)
    ms = timeit(m, 10)
    print(f"{tag} ising2d {shape}: {ms:.3f} ms/sweep  {m.nall()/ms/1e6:.1f} flips/ns", flush=True)
    del m
m = ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42)
ms = timeit(m, 20)
print(f"{tag} ising3d: {ms:.3f} ms/sweep  {m.nall()/ms/1e6:.1f} flips/ns", flush=True)
