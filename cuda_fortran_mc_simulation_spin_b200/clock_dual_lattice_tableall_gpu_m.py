"""Host mirror of ``module clock_dual_lattice_tableall_gpu_m``
(src/clock/clock_dual_lattice_tableall_m.f90): the tableall model stored as two compact
colour arrays ``sixclock_even / sixclock_odd (nx/2, ny)``.  Same public procedures (:43-45) and
-- because the reference indexes its randoms by the full-lattice coordinate (:144-152) -- the
same trajectory as ``clock_tableall_gpu_m``; the B200 library stores exactly this layout (as
bytes), so both mirrors drive the same C-ABI object.
"""
from __future__ import annotations

import sys

from ._sixclock import sixclock as _sixclock

version = "GPU_dual_lattice_tableall"
mstate = 6
nx = 1000
ny = 1000
nall = nx * ny
kbt = 0.91
beta = 1 / kbt
n_multi = 1
clock_gpu_stat = 0

_state = None


def configure(nx_=None, ny_=None, kbt_=None, mstate_=None, n_multi_=None):
    global nx, ny, nall, kbt, beta, mstate, n_multi
    if nx_ is not None: nx = int(nx_)
    if ny_ is not None: ny = int(ny_)
    if kbt_ is not None: kbt = float(kbt_)
    if mstate_ is not None: mstate = int(mstate_)
    if n_multi_ is not None: n_multi = int(n_multi_)
    nall = nx * ny
    beta = 1 / kbt


def print_version():
    sys.stdout.write("#" + version + "\n")
    sys.stderr.write("#" + version + "\n")


def init_sixclock(iseed):  # :57-90
    global _state
    if _state is not None:
        _state.close()
    _state = _sixclock(nx, ny, kbt, mstate, n_multi, iseed)


def handle() -> _sixclock:
    if _state is None:
        raise RuntimeError("call init_sixclock(iseed) first")
    return _state


def skip_curand_clock(n_skip):
    if int(n_skip) != 0:
        handle().skip_curand_clock(n_skip)


def init_sixclock_order():  # :92-95
    handle().init_sixclock_order()


def update_metropolis():  # :96-104
    handle().update_metropolis()


def _scalar_or_array(a):
    return float(a[0]) if n_multi == 1 else a


def calc_magne():  # :158-173
    return _scalar_or_array(handle().calc_magne())


def calc_energy():  # :175-201
    return _scalar_or_array(handle().calc_energy())


def sixclock_even_odd():
    ev, od = handle().get_dual()
    return (ev[0], od[0]) if n_multi == 1 else (ev, od)
