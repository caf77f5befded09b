"""B200-native checkerboard lattice-spin Monte Carlo (hot path of
osada-yum/CUDA_Fortran_MC_simulation_spin) behind the reference's module API.

Host-side mirrors of the reference's Fortran modules (same module, type and
procedure names) over the C ABI of ``libb200mc.so``:

    ising2d_gpu_m.ising2d_gpu        src/ising2d_gpu_m.f90
    ising3d_gpu_m.ising3d_gpu        src/ising3d_gpu_m.f90
    clock_gpu_m.clock_gpu            src/clock_gpu_m.f90
    clock_gpu_multi_m.clock_gpu      src/clock_gpu_multi_m.f90
    clock_tableall_gpu_m             src/clock/clock_tableall_gpu_m.f90
    clock_dual_lattice_tableall_gpu_m  src/clock/clock_dual_lattice_tableall_m.f90
    clock_table_gpu_m, clock_simple_gpu_m  src/clock/clock_table_gpu_m.f90, clock_simple_gpu_m.f90
    xy2d_periodic_gpu_m.xy2d_gpu     src/xy2d_periodic_gpu_m.f90
    xy2d_gpu_m.xy2d_gpu              src/xy2d_gpu_m.f90 (helical boundary)
    ising_periodic_gpu_m.ising_periodic_gpu   (no reference module: Ising 2D / 3D on the torus, L = 1024^3)
"""
from ._lib import B200MCError, SO_PATH, build  # noqa: F401
