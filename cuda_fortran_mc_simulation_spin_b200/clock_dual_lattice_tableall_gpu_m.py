"""Host mirror of ``module clock_dual_lattice_tableall_gpu_m`` (src/clock/clock_dual_lattice_tableall_m.f90: the tableall model stored as two compact colour arrays sixclock_even / sixclock_odd (nx/2, ny); its randoms are indexed by the full-lattice coordinate (:144-152), so the trajectory equals tableall's -- the B200 library stores exactly this layout (as bytes)).
Public procedures and parameters as in the reference (module-level); see _sixclock_module.py."""
from ._sixclock_module import install as _install

_install(globals(), "GPU_dual_lattice_tableall", 1000, 0)
