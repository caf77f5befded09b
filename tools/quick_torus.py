"""Periodic (torus) Ising: ms per MCS (device-timed update_n), per-launch time of the colour pass (the library's event pairs)
and the drivers' loop update + E + M per MCS, next to the helical module at the neighbouring shape."""
import os
import sys
import time

sys.path.insert(0, ".")
import torch

from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m, ising3d_gpu_m
from cuda_fortran_mc_simulation_spin_b200 import ising_periodic_gpu_m as M

PEAK = 6456.5


def run(name, m, n):
    m.update_n(5); m.sync()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        m.update_n(n)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n)
    m.set_timing(True)
    m.update_n(n); m.sync()
    n_pass, pass_ms = m.get_timing()
    m.set_timing(False)
    m.update(); m.measure(); m.update(); m.measure()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        m.update(); m.measure()
    torch.cuda.synchronize(); e2e = (time.perf_counter() - t0) * 1e3 / n
    nall = m.nall()
    med = sorted(ts)[len(ts) // 2]
    k_us = pass_ms / max(n_pass, 1) * 1e3
    print(f"{name}: update_n min {min(ts):.4f} med {med:.4f} ms/MCS = {nall / med / 1e6:.0f} flips/ns ({nall / med / 1e6 * 3 / PEAK * 100:.1f} % of HBM roofline); "
          f"pass kernel {k_us:.1f} us = {nall / 2 * 3 / (k_us * 1e-6) / 1e9 / PEAK * 100:.1f} %; loop update+E+M {e2e:.4f} ms/MCS = {nall / e2e / 1e6:.0f} flips/ns",
          flush=True)


tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("B200MC_"))
print("env:", tag or "(default)", flush=True)
run("torus 3d 1024^3", M.ising_periodic_gpu().init(1024, 1024, 1024, 4.51152, 42), 40)
run("helical 3d 1023x1023x1024", ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42), 40)
run("torus 2d 65536^2", M.ising_periodic_gpu().init(65536, 65536, 0, 2.26918531421, 42), 12)
run("helical 2d 65537x65536", ising2d_gpu_m.ising2d_gpu().init(65537, 65536, 2.26918531421, 42), 12)
run("torus 2d 16384^2", M.ising_periodic_gpu().init(16384, 16384, 0, 2.26918531421, 42), 100)
run("torus 3d 2048x512x512", M.ising_periodic_gpu().init(2048, 512, 512, 4.51152, 42), 60)
