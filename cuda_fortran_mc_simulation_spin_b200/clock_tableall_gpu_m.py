"""Host mirror of ``module clock_tableall_gpu_m`` (src/clock/clock_tableall_gpu_m.f90).

The reference module keeps its state in module variables and bakes ``nx, ny, kbt, mstate``
in as compile-time parameters that ``scripts/fpm_run_clock_simple_core.sh:71-74`` patches with
``sed`` before every build (:10-15).  Here they are module attributes set by ``configure``
(no rebuild); the public procedures keep the reference's names (:43-45):
``init_sixclock, skip_curand_clock, init_sixclock_order, update_metropolis, calc_energy,
calc_magne, print_version``.  ``n_multi > 1`` runs that many independent samples at once
(the drivers' ``tot_sample`` loop batched); ``calc_*`` then return arrays.
"""
from __future__ import annotations

import sys

from ._sixclock import sixclock as _sixclock

version = "GPU_tableall"
mstate = 6
nx = 2000
ny = 2000
nall = nx * ny
kbt = 0.91
beta = 1 / kbt
n_multi = 1
clock_gpu_stat = 0

_state = None


def configure(nx_=None, ny_=None, kbt_=None, mstate_=None, n_multi_=None):
    """replaces the sed-patching of the module parameters (run before init_sixclock)"""
    global nx, ny, nall, kbt, beta, mstate, n_multi
    if nx_ is not None: nx = int(nx_)
    if ny_ is not None: ny = int(ny_)
    if kbt_ is not None: kbt = float(kbt_)
    if mstate_ is not None: mstate = int(mstate_)
    if n_multi_ is not None: n_multi = int(n_multi_)
    nall = nx * ny
    beta = 1 / kbt


def print_version():  # :46-49
    sys.stdout.write("#" + version + "\n")
    sys.stderr.write("#" + version + "\n")


def init_sixclock(iseed):  # :57-88
    global _state
    if _state is not None:
        _state.close()
    _state = _sixclock(nx, ny, kbt, mstate, n_multi, iseed)


def handle() -> _sixclock:
    if _state is None:
        raise RuntimeError("call init_sixclock(iseed) first")
    return _state


def skip_curand_clock(n_skip):  # :51-55
    if int(n_skip) != 0:
        handle().skip_curand_clock(n_skip)


def init_sixclock_order():  # :90-92
    handle().init_sixclock_order()


def update_metropolis():  # :94-102
    handle().update_metropolis()


def _scalar_or_array(a):
    return float(a[0]) if n_multi == 1 else a


def calc_magne():  # :155-165
    return _scalar_or_array(handle().calc_magne())


def calc_energy():  # :167-181
    return _scalar_or_array(handle().calc_energy())


def sixclock():
    """the module's device array sixclock(nx, ny) copied to the host (Fortran order)"""
    a = handle().get_sixclock()
    return a[0] if n_multi == 1 else a
