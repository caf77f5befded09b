"""time the bit-packed Ising 3D / 2D colour pass at the BASELINE sizes (events around every launch)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m, ising3d_gpu_m
def run(g, label, nsweep=20):
    g.update_n(3); g.sync()
    g.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.update_n(nsweep); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nsweep
    nl, tot = g.get_timing(); g.set_timing(False)
    n = g.nall()
    print(f"{label}: {ms:.3f} ms/MCS  {n/ms/1e6:.0f} flips/ns  (kernel {tot/nl:.4f} ms/launch = {n/2/(tot/nl)/1e6:.0f} flips/ns; "
          f"{3/8*n/2/(tot/nl)/1e6:.0f} GB/s of 3/8 B per flip)  E, M = {g.measure()}", flush=True)
g = ising3d_gpu_m.ising3d_gpu().init_packed(1023, 1023, 1024, 4.51152, 42)
run(g, "bits 3D 1023x1023x1024 kbt=4.51152 all-up start")
g.set_random_spin(); run(g, "bits 3D 1023x1023x1024 from disorder")
del g
g = ising2d_gpu_m.ising2d_gpu().init_packed(65537, 65536, 2.26918531421, 42)
run(g, "bits 2D 65537x65536 kbt=Tc all-up start", 10)
del g
g = ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42)
run(g, "int8 3D 1023x1023x1024 (for comparison)")
