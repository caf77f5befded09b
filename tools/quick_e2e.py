"""update + measure every MCS (the drivers' loop), fused vs separate measurement (B200MC_TUNE=8 disables fusing)"""
import sys, os
sys.path.insert(0, ".")
import torch
from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m, ising2d_gpu_m
def loop(m, n):
    acc = 0
    for _ in range(n):
        m.update(); acc += m.calc_magne_sum(); acc += m.calc_energy_sum()
    return acc
def timeit(m, n):
    loop(m, 3); m.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    loop(m, n)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tag = f"TUNE={os.environ.get('B200MC_TUNE')}"
m = ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42)
ms = timeit(m, 20); print(f"{tag} ising3d update+measure: {ms:.3f} ms/MCS  {m.nall()/ms/1e6:.1f} flips/ns", flush=True)
m.update_n(3); m.sync(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.update_n(20); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20; print(f"{tag} ising3d update only:    {ms:.3f} ms/MCS  {m.nall()/ms/1e6:.1f} flips/ns", flush=True)
del m
m = ising2d_gpu_m.ising2d_gpu().init(1001, 1000, 2.26918531421, 42)
ms = timeit(m, 500); print(f"{tag} ising2d 1001x1000 update+measure: {ms*1e3:.1f} us/MCS  {m.nall()/ms/1e6:.1f} flips/ns", flush=True)
