import sys
sys.path.insert(0, ".")
from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m
m = ising3d_gpu_m.ising3d_gpu().init(1023, 1023, 1024, 4.51152, 42)
m.update_n(int(sys.argv[1]) if len(sys.argv) > 1 else 4)
print(m.measure())
