"""Builds the module-procedure API shared by the reference's periodic-clock modules (src/clock/*.f90):
``init_sixclock, skip_curand_clock, init_sixclock_order, update_metropolis, calc_energy, calc_magne,
print_version`` plus the public parameters ``mstate, nx, ny, nall, kbt, beta`` -- which the reference bakes in at
compile time and patches with ``sed`` (scripts/fpm_run_clock_simple_core.sh:71-74); ``configure`` replaces that."""
from __future__ import annotations

import sys

from ._sixclock import sixclock as _sixclock


def install(ns, version, default_nx, variant):
    ns.update(version=version, mstate=6, nx=default_nx, ny=default_nx, nall=default_nx * default_nx, kbt=0.91,
              beta=1 / 0.91, n_multi=1, clock_gpu_stat=0, _state=None)

    def configure(nx_=None, ny_=None, kbt_=None, mstate_=None, n_multi_=None):
        """replaces the sed-patching of the module parameters (run before init_sixclock)"""
        if nx_ is not None: ns["nx"] = int(nx_)
        if ny_ is not None: ns["ny"] = int(ny_)
        if kbt_ is not None: ns["kbt"] = float(kbt_)
        if mstate_ is not None: ns["mstate"] = int(mstate_)
        if n_multi_ is not None: ns["n_multi"] = int(n_multi_)
        ns["nall"] = ns["nx"] * ns["ny"]
        ns["beta"] = 1 / ns["kbt"]

    def print_version():
        sys.stdout.write("#" + version + "\n")
        sys.stderr.write("#" + version + "\n")

    def init_sixclock(iseed):
        if ns["_state"] is not None:
            ns["_state"].close()
        ns["_state"] = _sixclock(ns["nx"], ns["ny"], ns["kbt"], ns["mstate"], ns["n_multi"], iseed, variant=variant)

    def handle():
        if ns["_state"] is None:
            raise RuntimeError("call init_sixclock(iseed) first")
        return ns["_state"]

    def skip_curand_clock(n_skip):
        if int(n_skip) != 0:
            handle().skip_curand_clock(n_skip)

    def init_sixclock_order():
        handle().init_sixclock_order()

    def update_metropolis():
        handle().update_metropolis()

    def _scalar_or_array(a):
        return float(a[0]) if ns["n_multi"] == 1 else a

    def calc_magne():
        return _scalar_or_array(handle().calc_magne())

    def calc_energy():
        return _scalar_or_array(handle().calc_energy())

    def sixclock():
        """the module's device array sixclock(nx, ny) copied to the host (Fortran order)"""
        a = handle().get_sixclock()
        return a[0] if ns["n_multi"] == 1 else a

    def sixclock_even_odd():
        ev, od = handle().get_dual()
        return (ev[0], od[0]) if ns["n_multi"] == 1 else (ev, od)

    for f in (configure, print_version, init_sixclock, handle, skip_curand_clock, init_sixclock_order, update_metropolis,
              calc_magne, calc_energy, sixclock, sixclock_even_odd):
        ns[f.__name__] = f
