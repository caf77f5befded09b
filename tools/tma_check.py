"""B200MC_TUNE=128: copy-engine staged Ising pass (ising_pass_tma_kernel); parity vs oracle on a lattice large enough
for the ticket path + timing at the headline size"""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m, ising2d_gpu_m
tag = f"TUNE={os.environ.get('B200MC_TUNE')}"
if "check" in sys.argv:
    from oracle import oracle as O
    for shape, method in [((255, 257, 320), 0), ((255, 257, 322), 1)]:
        g = ising3d_gpu_m.ising3d_gpu().init(*shape, 4.51152, 42); o = O.ising3d_gpu().init(*shape, 4.51152, 42)
        g.set_method(method)
        g.set_random_spin(); o.set_random_spin()
        for i in range(3):
            g.update(); (o.update_heatbath if method else o.update)()
            a, b = g.spins(), o.spins()
            if not np.array_equal(a, b):
                bad = np.nonzero(a != b)[0]
                print(tag, shape, "sweep", i, "MISMATCH", bad.size, bad[:6], flush=True); break
            em = g.measure()
            assert em == (o.calc_energy_sum(), o.calc_magne_sum()), (shape, i, em)
        else:
            print(tag, shape, "method", method, "parity ok (spins, fused E/M)", flush=True)
    g = ising2d_gpu_m.ising2d_gpu().init(4097, 5120, 2.269, 42); o = O.ising2d_gpu().init(4097, 5120, 2.269, 42)
    for i in range(3):
        g.update(); o.update()
        assert np.array_equal(g.spins(), o.spins()), i
        assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())
    print(tag, "2d parity ok", flush=True)
m = ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42)
m.update_n(5); m.sync()
m.set_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); m.update_n(30); e1.record(); torch.cuda.synchronize()
n, ms = m.get_timing()
print(f"{tag} ising3d: {e0.elapsed_time(e1)/30:.4f} ms/MCS  {m.nall()*30/e0.elapsed_time(e1)/1e6:.1f} flips/ns  pass kernel {ms/n*1e3:.1f} us", flush=True)
del m
m = ising2d_gpu_m.ising2d_gpu().init(65537, 65536, 2.269, 42)
m.update_n(3); m.sync(); m.set_timing(True)
torch.cuda.synchronize(); e0.record(); m.update_n(10); e1.record(); torch.cuda.synchronize()
n, ms = m.get_timing()
print(f"{tag} ising2d: {e0.elapsed_time(e1)/10:.4f} ms/MCS  {m.nall()*10/e0.elapsed_time(e1)/1e6:.1f} flips/ns  pass kernel {ms/n*1e3:.1f} us", flush=True)
