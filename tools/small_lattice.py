"""Small-lattice regime (BASELINE config 1: the reference's default 1001 x 1000): per-MCS cost of update_n, of the
drivers' loop update -> calc_magne_sum -> calc_energy_sum, and of run_relaxation, with the cooperative sweep kernel
and with the launch-per-pass path (B200MC_TUNE bit 11), and that both give the same trajectory."""
import os, sys, time
sys.path.insert(0, ".")
import torch
from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m, ising3d_gpu_m

def run(tune, kind, shape, kbt):
    os.environ["B200MC_TUNE"] = str(tune)
    mod = ising2d_gpu_m.ising2d_gpu if kind == "2d" else ising3d_gpu_m.ising3d_gpu
    m = mod().init(*shape, kbt, 42)
    m.update_n(10); m.sync()
    torch.cuda.synchronize(); t0 = time.perf_counter(); m.update_n(500); m.sync(); t_upd = (time.perf_counter() - t0) / 500
    series = []
    for _ in range(5): m.update(); series.append((m.calc_magne_sum(), m.calc_energy_sum()))
    t0 = time.perf_counter()
    for _ in range(300): m.update(); series.append((m.calc_magne_sum(), m.calc_energy_sum()))
    t_loop = (time.perf_counter() - t0) / 300
    m.run_relaxation(10)
    t0 = time.perf_counter(); e, mg = m.run_relaxation(1000); t_rel = (time.perf_counter() - t0) / 1000
    n = m.nall()
    print(f"{kind} {shape} TUNE={tune}: update_n {t_upd*1e6:.1f} us/MCS ({n/t_upd/1e9:.1f} flips/ns)  driver loop {t_loop*1e6:.1f} us/MCS ({n/t_loop/1e9:.1f})  "
          f"run_relaxation {t_rel*1e6:.1f} us/MCS ({n/t_rel/1e9:.1f})", flush=True)
    return series, list(map(int, e[-3:])), list(map(int, mg[-3:])), m.measure()

for kind, shape, kbt in (("2d", (1001, 1000), 2.26918531421), ("2d", (1025, 1024), 2.26918531421), ("3d", (101, 101, 100), 4.51152), ("2d", (2049, 2048), 2.26918531421)):
    a = run(0, kind, shape, kbt)
    b = run(2048, kind, shape, kbt)
    assert a == b, "cooperative path differs from the launch-per-pass path"
print("same trajectories")
