"""A/B a library build: pass-kernel time for 3D plain, 3D fused-measure loop, 3D self-push (TUNE=16 set by caller), 2D"""
import os, sys
sys.path.insert(0, ".")
import torch
from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m, ising2d_gpu_m
def ev(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
m = ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42)
m.update_n(5); m.sync()
t3 = ev(m.update_n, 30)
def loop(n):
    for _ in range(n):
        m.update(); m.calc_magne_sum(); m.calc_energy_sum()
loop(3); te = ev(loop, 20)
del m
m = ising2d_gpu_m.ising2d_gpu().init(65537, 65536, 2.269, 42)
m.update_n(3); m.sync(); t2 = ev(m.update_n, 10)
print(f"{os.path.basename(os.environ.get('B200MC_SO','default'))} TUNE={os.environ.get('B200MC_TUNE')}: 3D {t3:.4f} ms/MCS ({1071645696/t3/1e6:.0f} flips/ns)  3D+measure {te:.4f} ms ({1071645696/te/1e6:.0f})  2D {t2:.4f} ms ({m.nall()/t2/1e6:.0f})", flush=True)
