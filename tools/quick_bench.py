"""scratch timing of the Ising sweeps (not the contract bench; see bench.py)"""
import sys, time
sys.path.insert(0, ".")
import torch
from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m, ising2d_gpu_m

def timeit(m, n):
    m.update_n(3); m.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    m.update_n(n)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return ms / n

for shape in [(1023, 1023, 1024), (511, 511, 512)]:
    m = ising3d_gpu_m.ising3d_gpu().init(*shape, 4.51152, 42)
    ms = timeit(m, 20)
    print(f"ising3d {shape}: {ms:.3f} ms/sweep  {m.nall()/ms/1e6:.1f} flips/ns  E,M={m.measure()}", flush=True)
    m.set_method(1)
    ms = timeit(m, 20)
    print(f"  heatbath: {ms:.3f} ms/sweep  {m.nall()/ms/1e6:.1f} flips/ns", flush=True)
    del m
for shape in [(1001, 1000), (16385, 16384), (65537, 65536)]:
    m = ising2d_gpu_m.ising2d_gpu().init(*shape, 2.26918531421, 42)
    ms = timeit(m, 20)
    print(f"ising2d {shape}: {ms:.4f} ms/sweep  {m.nall()/ms/1e6:.1f} flips/ns  E,M={m.measure()}", flush=True)
    del m
