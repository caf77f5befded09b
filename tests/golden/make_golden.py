#!/usr/bin/env python
"""Generates tests/golden/trajectories.json.

PROVENANCE.  The reference (CUDA Fortran + OpenACC + cuRAND) holds no golden vectors and cannot
be built or run in this image (no Fortran compiler), so these vectors are outputs of the CPU
oracle (oracle/oracle.c: the restatement of the reference kernels, file:line cited there) fed with
the RNG contract's uniforms (oracle/rng_contract.c).  They pin (i) the oracle itself against
accidental change (tests/test_golden.py, CPU) and (ii) the CUDA path on the GPU box without the
oracle in the loop (tests/test_gpu_golden.py).  Regenerate with:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

KBT3, KBT2 = 4.51152, 2.26918531421


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ising_cases(O):
    out = []
    specs = [("ising3d", (31, 31, 30), KBT3, 42, "allup", 0, 6), ("ising3d", (63, 65, 64), KBT3, 7, "random", 0, 5),
             ("ising3d", (63, 65, 64), KBT3, 7, "random", 1, 4), ("ising3d", (5, 5, 4), KBT3, 1, "random", 0, 8),
             ("ising2d", (1001, 1000), KBT2, 42, "allup", 0, 4), ("ising2d", (255, 256), KBT2, 3, "random", 0, 5),
             ("ising2d", (255, 256), KBT2, 3, "random", 1, 4), ("ising2d", (5, 4), KBT2, 9, "random", 0, 8)]
    for model, shape, kbt, seed, start, method, sweeps in specs:
        o = (O.ising3d_gpu() if model == "ising3d" else O.ising2d_gpu()).init(*shape, kbt, seed)
        if start == "random":
            o.set_random_spin()
        series = []
        for _ in range(sweeps):
            (o.update_heatbath if method else o.update)()
            series.append([o.calc_energy_sum(), o.calc_magne_sum()])
        out.append({"model": model, "shape": list(shape), "kbt": kbt, "seed": seed, "start": start, "method": method,
                    "em": series, "spins_sha256": sha(o.spins()), "spins_head": o.spins()[:24].tolist()})
    return out


def torus_cases(O):
    """periodic (torus) Ising: nz = 0 -> 2D; nx = 1024 rows run the strip kernel on the GPU, the others the generic one"""
    out = []
    specs = [((64, 6, 4), KBT3, 42, "allup", 0, 6), ((1024, 8, 4), KBT3, 7, "random", 0, 5), ((1024, 6, 2), KBT3, 7, "random", 1, 4),
             ((96, 10, 0), KBT2, 3, "random", 0, 5), ((2048, 8, 0), KBT2, 9, "allup", 1, 4)]
    for shape, kbt, seed, start, method, sweeps in specs:
        o = O.ising_periodic_gpu().init(*shape, kbt, seed)
        if start == "random":
            o.set_random_spin()
        series = []
        for _ in range(sweeps):
            (o.update_heatbath if method else o.update)()
            series.append(list(o.measure()))
        out.append({"model": "ising_torus", "shape": list(shape), "kbt": kbt, "seed": seed, "start": start, "method": method,
                    "em": series, "spins_sha256": sha(o.spins()), "spins_head": o.spins()[:24].tolist()})
    return out


def clock_cases(O):
    out = []
    for shape, q, kbt, seed, n_multi, start, sweeps in [((33, 32), 6, 0.91, 42, None, "allup", 5), ((101, 100), 6, 0.8, 5, None, "random", 4),
                                                        ((31, 30), 5, 0.9, 2, None, "random", 4), ((65, 64), 6, 0.8, 42, 3, "random", 3)]:
        o = O.clock_gpu().init(*shape, kbt, q, seed, n_multi=n_multi)
        if start == "random":
            o.set_random_spin()
        hists = []
        for _ in range(sweeps):
            o.update()
            hists.append([o.histograms(j)[0].tolist() for j in range(o.n_multi_)])
        out.append({"model": "clock", "shape": list(shape), "q": q, "kbt": kbt, "seed": seed, "n_multi": n_multi, "start": start,
                    "hist": hists, "spins_sha256": sha(o.spins()),
                    "energy": np.atleast_1d(o.calc_energy_sum()).tolist(), "magne": np.atleast_1d(o.calc_magne_sum()).tolist()})
    return out


def sixclock_cases(O):
    out = []
    for shape, q, kbt, seed, n_multi, sweeps in [((64, 64), 6, 0.91, 42, 1, 5), ((200, 100), 6, 0.91, 3, 1, 4), ((34, 6), 6, 0.8, 1, 2, 6),
                                                 ((96, 12), 5, 0.9, 8, 1, 4)]:
        nx, ny = shape
        os_ = [O.clock_tableall(nx, ny, kbt, q) for _ in range(n_multi)]
        hists = []
        for s in range(sweeps):
            for j, o in enumerate(os_):
                o.update_metropolis(O.torus_uniforms(seed, s, j, nx, ny, q))
            hists.append([o.histograms()[0].tolist() for o in os_])
        out.append({"model": "sixclock", "shape": list(shape), "q": q, "kbt": kbt, "seed": seed, "n_multi": n_multi,
                    "hist": hists, "states_sha256": sha(np.stack([o.c for o in os_])),
                    "energy": [o.calc_energy() for o in os_], "magne": [o.calc_magne() for o in os_]})
    return out


def xy_cases(O):
    """one Metropolis sweep + one over-relaxation step from the contract's random start, real64 oracle"""
    out = []
    for shape, kbt, seed in [((64, 64), 0.89, 42), ((256, 128), 0.5, 7)]:
        nx, ny = shape
        o = O.xy2d_gpu().init(nx, ny, kbt, seed)
        o.set_random_spin(O.xy_init_uniforms(seed, 0, nx, ny))
        start = [o.calc_energy_sum(), o.calc_magne_sum(), o.calc_magne_y_sum()]
        r, c = O.xy_uniforms(seed, 1, nx, ny)
        o.update(r, c)
        after_m = [o.calc_energy_sum(), o.calc_magne_sum(), o.calc_magne_y_sum()]
        o.update_over_relaxation(1)
        after_o = [o.calc_energy_sum(), o.calc_magne_sum(), o.calc_magne_y_sum()]
        out.append({"model": "xy2d", "shape": list(shape), "kbt": kbt, "seed": seed, "start": start,
                    "after_metropolis": after_m, "after_over_relaxation": after_o})
    return out


def rng_cases(O):
    return {"ising_uniforms(42,3,64)": O.ising_uniforms(42, 3, 64).tolist(),
            "ring_init_uniforms(42,0,32)": O.ring_init_uniforms(42, 0, 32).tolist(),
            "clock_uniforms(42,2,1,32)": [a.tolist() for a in O.clock_uniforms(42, 2, 1, 32)],
            "torus_uniforms(42,2,1,8,4)": O.torus_uniforms(42, 2, 1, 8, 4).tolist(),
            "xy_uniforms(42,2,8,4)": [a.tolist() for a in O.xy_uniforms(42, 2, 8, 4)],
            "isingp_uniforms(42,3,32,2,2)": O.isingp_uniforms(42, 3, 32, 2, 2).tolist()}


def generate():
    from oracle import oracle as O
    O.build()
    return {"provenance": "CPU oracle (oracle/oracle.c + oracle/rng_contract.c); the reference has no golden vectors and cannot run here",
            "ising": ising_cases(O), "ising_torus": torus_cases(O), "clock": clock_cases(O), "sixclock": sixclock_cases(O), "xy": xy_cases(O), "rng": rng_cases(O)}


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "trajectories.json")
    with open(path, "w") as f:
        json.dump(generate(), f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")
