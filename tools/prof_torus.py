"""a few sweeps of one model for ncu: prof_torus.py torus3d | helical3d | torus2d [fused]"""
import sys
sys.path.insert(0, ".")
from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m
from cuda_fortran_mc_simulation_spin_b200 import ising_periodic_gpu_m as M
which = sys.argv[1]
if which == "torus3d":
    m = M.ising_periodic_gpu().init(1024, 1024, 1024, 4.51152, 42)
elif which == "torus2d":
    m = M.ising_periodic_gpu().init(65536, 65536, 0, 2.26918531421, 42)
else:
    m = ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42)
m.update_n(20)   # away from the all-up start: the tie / accept rates of the equilibrium regime
m.sync()
if len(sys.argv) > 2:
    for _ in range(4):
        m.update(); m.measure()
else:
    m.update_n(4)
m.sync()
