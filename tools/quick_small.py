"""C1: Ising 2D at the reference's default 1001 x 1000 (app/ising2d_gpu_relaxation.f90:6-12): per-MCS host loop vs
the device-side driver loop"""
import sys, time
sys.path.insert(0, ".")
from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m, ising3d_gpu_m
for shape in [(1001, 1000), (1023, 1024)]:
    g = ising2d_gpu_m.ising2d_gpu().init(*shape, 2.26918531421, 42)
    g.run_relaxation(50); n = 1000
    t0 = time.perf_counter(); e, m = g.run_relaxation(n); dt = time.perf_counter() - t0
    print(f"ising2d {shape} run_relaxation: {dt/n*1e6:.1f} us/MCS  {g.nall()*n/dt/1e9:.1f} flips/ns  m(1000)={m[-1]/g.nall():.4f}", flush=True)
    t0 = time.perf_counter()
    for _ in range(n):
        g.update(); g.calc_magne_sum(); g.calc_energy_sum()
    dt = time.perf_counter() - t0
    print(f"ising2d {shape} host loop:      {dt/n*1e6:.1f} us/MCS  {g.nall()*n/dt/1e9:.1f} flips/ns", flush=True)
g = ising3d_gpu_m.ising3d_gpu().init(101, 101, 100, 4.51152, 42)
g.run_relaxation(50); n = 1000
t0 = time.perf_counter(); e, m = g.run_relaxation(n); dt = time.perf_counter() - t0
print(f"ising3d 101x101x100 run_relaxation: {dt/n*1e6:.1f} us/MCS  {g.nall()*n/dt/1e9:.1f} flips/ns", flush=True)
for nm in (8, 32, 128):
    g = ising2d_gpu_m.ising2d_gpu().init_multi(1001, 1000, 2.26918531421, 42, nm)
    g.run_relaxation(20); n = 200
    t0 = time.perf_counter(); e, m = g.run_relaxation(n); dt = time.perf_counter() - t0
    print(f"ising2d 1001x1000 x {nm} samples run_relaxation: {dt/n*1e6:.1f} us/MCS  {g.nall()*nm*n/dt/1e9:.1f} flips/ns  <m(200)>={m[:, -1].mean()/g.nall():.4f}", flush=True)
