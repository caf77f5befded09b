"""Host mirror of ``module clock_table_gpu_m`` (src/clock/clock_table_gpu_m.f90: delta-E from four lookups of the q^3 energy table and exp() per site (:119-125) -- the same expression as tableall's table entries, so the same decisions).
Public procedures and parameters as in the reference (module-level); see _sixclock_module.py."""
from ._sixclock_module import install as _install

_install(globals(), "GPU_table", 1000, 0)
