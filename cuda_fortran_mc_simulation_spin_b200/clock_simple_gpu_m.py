"""Host mirror of ``module clock_simple_gpu_m`` (src/clock/clock_simple_gpu_m.f90: no tables, delta-E summed over the four neighbours per site (:108-113); the library tabulates exactly that expression).
Public procedures and parameters as in the reference (module-level); see _sixclock_module.py."""
from ._sixclock_module import install as _install

_install(globals(), "GPU_simple", 1000, 1)
