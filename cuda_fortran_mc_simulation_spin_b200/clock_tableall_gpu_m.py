"""Host mirror of ``module clock_tableall_gpu_m`` (src/clock/clock_tableall_gpu_m.f90: full q^6 acceptance table states_to_prob (:66-86)).
Public procedures and parameters as in the reference (module-level); see _sixclock_module.py."""
from ._sixclock_module import install as _install

_install(globals(), "GPU_tableall", 2000, 0)
