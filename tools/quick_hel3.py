"""helical 3D 1023x1023x1024 + 2D 65537x65536: update_n and loop -- one line per build (B200MC_SO)"""
import os, sys, time
sys.path.insert(0, ".")
import torch
from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m, ising3d_gpu_m
def run(name, m, n):
    m.update_n(5); m.sync()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); m.update_n(n); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n)
    m.update(); m.measure(); m.update(); m.measure()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        m.update(); m.measure()
    torch.cuda.synchronize(); e2e = (time.perf_counter() - t0) * 1e3 / n
    nall = m.nall(); med = sorted(ts)[len(ts) // 2]
    print(f"{os.path.basename(os.environ.get('B200MC_SO') or 'default'):28s} {name}: update_n min {min(ts):.4f} med {med:.4f} ms/MCS = {nall / med / 1e6:.0f} flips/ns ({nall / med / 1e6 * 3 / 6456.5 * 100:.1f} %); loop {e2e:.4f} ms = {nall / e2e / 1e6:.0f} flips/ns; E,M={m.measure()}", flush=True)
run("helical 3d", ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42), 40)
run("helical 2d", ising2d_gpu_m.ising2d_gpu().init(65537, 65536, 2.26918531421, 42), 12)
