"""B200MC_TUNE=16: single GPU through the fused update + halo kernel (self-neighbour); parity vs oracle + timing"""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m, ising2d_gpu_m
tag = f"TUNE={os.environ.get('B200MC_TUNE')}"
if "check" in sys.argv:
    from oracle import oracle as O
    for shape in [(63, 65, 128), (31, 31, 256)]:
        g = ising3d_gpu_m.ising3d_gpu().init(*shape, 4.51152, 42); o = O.ising3d_gpu().init(*shape, 4.51152, 42)
        g.set_random_spin(); o.set_random_spin()
        for i in range(5):
            g.update(); o.update()
            assert np.array_equal(g.spins(), o.spins()), (shape, i)
            assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())
    g = ising2d_gpu_m.ising2d_gpu().init(255, 1024, 2.269, 42); o = O.ising2d_gpu().init(255, 1024, 2.269, 42)
    for i in range(5):
        g.update(); o.update()
        assert np.array_equal(g.spins(), o.spins()), i
        assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())
    print(tag, "parity ok", flush=True)
m = ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42)
m.update_n(5); m.sync()
m.set_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); m.update_n(30); e1.record(); torch.cuda.synchronize()
n, ms = m.get_timing()
print(f"{tag} ising3d: {e0.elapsed_time(e1)/30:.4f} ms/MCS  {m.nall()*30/e0.elapsed_time(e1)/1e6:.1f} flips/ns  pass kernel {ms/n*1e3:.1f} us", flush=True)
