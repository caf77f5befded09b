#!/usr/bin/env python
"""bench.py -- headline benchmark of the checkerboard lattice-spin MC sweep.

Metric (BASELINE.json): attempted spin flips / ns, device-timed, and the
fraction of the HBM roofline.  Headline workload at N GPUs (weak scaling):
BASELINE config 2 (C2), the Ising 3D checkerboard Metropolis relaxation on an
int8 lattice, at the reference-valid helical shape next to 1024^3:
1023 x 1023 x 1024 per GPU (1024^3 itself is rejected by the reference's
linear-index colouring), kbt = 4.51152, all-up start
(app/ising3d_gpu_relaxation.f90).  A "step" is one MCS = one `update()` =
nall attempted flips per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--only-headline]

Prints ONE JSON line (rank 0).  Besides the headline keys the line carries
`configs`: the other BASELINE configurations (C2 / C5 with true periodic boundaries -- L = 1024^3 itself --, C1 Ising 2D small lattice, C3 XY
16384^2 Metropolis / over-relaxation, C4 q=6 clock 16384^2 batched + helical,
C5 Ising 2D 65537 x 65536 per GPU), each device-timed the same way with its own
roofline fraction.  At N > 1 the sharded workloads run: C2 (headline) and C5.

Timing: every number is the median over `blocks` timed blocks of EXACTLY K steps,
each bracketed by barrier + synchronize and CUDA events, max over ranks; blocks
are repeated until >= 1 s has been spent under load so that the clock sampler
(and the driver's) sees the timed region.

`--impl reference` times the CPU restatement of the reference's own algorithm
(oracle/, kind "port": the reference is CUDA Fortran and cannot be built in this
image) on the host cores: all cores (the line's value) and one core.
"""
from __future__ import annotations

import argparse
import gc
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NX, NY, NZ = 1023, 1023, 1024
KBT = 4.51152
KBT2 = 2.26918531421
SEED = 42
METRIC = "spin_flips_per_ns"
UNIT = "flips/ns"
BYTES_PER_FLIP = 3.0  # int8 two-pass checkerboard: read other colour, read own, write own
BYTES_PER_FLIP_XY = 12.0  # fp32 angle per site, same three streams
CPU_SHAPE = (255, 255, 256)  # bounded CPU sample: same model / temperature / start, 1/64 of the sites
MIN_TIMED_SECONDS = 1.0


def _measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic():
    """DRAM bytes per launch of the pass kernel from the committed ncu capture (static: not re-measured by this run)."""
    for name in ("r02_ising3d_pass_traffic.json", "r01_ising3d_pass_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        try:
            with open(p) as f:
                return float(json.load(f)["dram_bytes_per_launch"]), f"static: profiles/{name} (ncu --set full capture of the same kernel and shape)"
        except Exception:
            continue
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(n)
        # "under load": samples drawing more than half of the largest power seen (the timed regions)
        pmax = max(power) if power else 0.0
        load = [s for s, p in zip(sm, power) if p >= 0.5 * pmax] or sm
        return {"sm_mhz": statistics.median(load) if load else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": pmax if power else None,
                "samples": len(sm), "samples_under_load": len(load), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference) on the host cores
# ----------------------------------------------------------------------------------------------
def _host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _oracle_sweep_rate(threads: int, steps: int, warmup: int, min_seconds: float = 0.0, max_seconds: float = 60.0):
    """Time the CPU restatement of the reference's per-MCS work on `threads` host threads:
    draw nall uniforms (the reference's curandGenerate, src/ising3d_gpu_m.f90:179), two
    colour passes + halo copies (:180-187), then the two reductions the drivers call every
    MCS (calc_magne_sum, calc_energy_sum; app/ising3d_gpu_relaxation.f90:43-45)."""
    import numpy as np

    from oracle import oracle as O

    O.build()
    O.set_threads(int(threads))   # explicit: torchrun exports OMP_NUM_THREADS=1
    nx, ny, nz = CPU_SHAPE
    m = O.ising3d_gpu().init(nx, ny, nz, KBT, SEED)
    u = np.empty(m.nall(), dtype=np.float64)

    def one():
        O.ising_uniforms_fast(m.seed_, m.draw_, m.nall(), out=u)
        m.draw_ += 1
        m.update(randoms=u)
        m.calc_magne_sum()
        m.calc_energy_sum()

    for _ in range(max(warmup, 1)):
        one()
    t0 = time.perf_counter()
    done = 0
    while (done < steps or (time.perf_counter() - t0) < min_seconds) and (time.perf_counter() - t0) < max_seconds:
        one()
        done += 1
    dt = time.perf_counter() - t0
    return {"flips_per_ns": m.nall() * done / dt / 1e9, "steps": done, "seconds": dt,
            "cores": int(threads), "nall": m.nall()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0   # one CPU arm per job: rank 0 times it on all host threads of the box, the other ranks do no work
    nthr = _host_threads()
    r = _oracle_sweep_rate(nthr, args.steps, args.warmup, max_seconds=120.0)
    r1 = _oracle_sweep_rate(1, 2, 1, max_seconds=30.0)
    sample = (f"Ising 3D helical {CPU_SHAPE[0]}x{CPU_SHAPE[1]}x{CPU_SHAPE[2]} (1/64 of the per-GPU lattice), "
              f"{r['steps']} MCS, per MCS: draw nall uniforms + 2 colour passes + halo copies + E and M reductions, "
              f"OpenMP on {r['cores']} host threads (all threads of the box, one process: rank 0)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["flips_per_ns"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["seconds"] / r["steps"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32 spins / f64 uniforms",
        "data": "synthetic",
        "config": {"workload": f"Ising 3D checkerboard Metropolis, helical {NX}x{NY}x{NZ} per GPU, kbt={KBT}, all-up start",
                   "note": "CPU port of the reference algorithm timed on a bounded sample; flips/ns is size-independent. "
                           "The box's host cores are one resource: the value does not grow with --gpus"},
        "cpu_baseline": {"value": r["flips_per_ns"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample,
                         "serial_value": r1["flips_per_ns"], "serial_cores": 1, "serial_steps": r1["steps"]},
        "e2e": {"value": r["flips_per_ns"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class Timer:
    """blocks of exactly K steps, CUDA events on the launching stream (the library launches on the legacy default
    stream = torch's current stream here), barrier + synchronize on both sides, max over ranks per block"""

    def __init__(self, torch, dist):
        self.torch, self.dist = torch, dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _max(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def block(self, fn):
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        fn()
        e1.record()
        self.barrier()
        return self._max(e0.elapsed_time(e1))

    def run(self, fn, min_seconds=MIN_TIMED_SECONDS, max_blocks=400):
        """fn() runs exactly K steps.  Returns the per-block times (ms, max over ranks)."""
        first = self.block(fn)
        n = int(min(max_blocks, max(2, math.ceil(min_seconds * 1e3 / max(first, 1e-3)))))   # same count on every rank: `first` is all-reduced
        return [first] + [self.block(fn) for _ in range(n)]


def _roof(flips_per_ns, bytes_per_flip, peak, basis):
    ach = flips_per_ns * bytes_per_flip
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "basis": basis}


def _kernel_roof(sites_per_launch, bytes_per_flip, n_launch, total_ms, peak, kernel):
    avg = total_ms / max(n_launch, 1)
    ach = bytes_per_flip * sites_per_launch / (avg * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "basis": "kernel",
            "kernel": kernel, "avg_launch_ms": avg, "launches_timed": n_launch}


def _free():
    import torch
    gc.collect()
    torch.cuda.synchronize()


def bench_other_configs(T: Timer, K: int, peak: float, world: int, rank: int):
    """BASELINE configs 1, 3, 4, 5 (5 only when world > 1), each: blocks of K steps, median block, roofline."""
    from cuda_fortran_mc_simulation_spin_b200 import clock_gpu_m, ising2d_gpu_m, xy2d_periodic_gpu_m
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock

    out = {}

    def entry(workload, sites_per_step, blocks, k, bpf, basis, extra=None):
        ms = statistics.median(blocks)
        v = sites_per_step * k / (ms * 1e6)
        e = {"workload": workload, "value": v, "unit": UNIT, "steps": k, "blocks": len(blocks), "ms_per_step": ms / k,
             "ms_per_step_min": min(blocks) / k, "roofline": _roof(v / world, bpf, peak, basis)}   # (per GPU)
        if extra:
            e.update(extra)
        return e

    # ---- C5: Ising 2D 65537 x 65536 per GPU, slabs along y (weak scaling) ----
    if world > 1:
        m = ising2d_gpu_m.ising2d_gpu().init_distributed(65537, 65536 * world, KBT2, SEED)
    else:
        m = ising2d_gpu_m.ising2d_gpu().init(65537, 65536, KBT2, SEED)
    nall = m.nall()
    k5 = max(2, K // 4)
    m.update_n(3); m.sync()
    m.set_timing(True)
    blocks = T.run(lambda: m.update_n(k5))
    n_pass, pass_ms = m.get_timing()
    m.set_timing(False)
    out["C5_ising2d_65537x65536_per_gpu"] = entry(
        f"Ising 2D Metropolis, int8, helical 65537x{65536 * world} in {world} slab(s) along y, kbt={KBT2}, all-up start",
        nall, blocks, k5, BYTES_PER_FLIP, "step",
        {"roofline_kernel": _kernel_roof(nall / world / 2, BYTES_PER_FLIP, n_pass, pass_ms, peak, "ising_pass_kernel<4, METROPOLIS>"),
         "sites_per_gpu": nall // world})
    del m
    _free()
    if world > 1:
        # ---- C2' in slabs: the bit-packed headline lattice, one slab per GPU (halos through NCCL send/recv, not overlapped) ----
        from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m
        m = ising3d_gpu_m.ising3d_gpu().init_packed_distributed(NX, NY, NZ * world, KBT, SEED)
        nall = m.nall()
        m.update_n(3); m.sync()
        blocks = T.run(lambda: m.update_n(K))
        out["C2p_ising3d_bitpacked_slabs"] = entry(
            f"Ising 3D Metropolis, ONE BIT per site, helical {NX}x{NY}x{NZ * world} in {world} slab(s), kbt={KBT}, all-up start",
            nall, blocks, K, 3.0 / 8.0, "step, against the bit-packed layout's own 3/8 B per flip", {"sites_per_gpu": nall // world})
        del m
        _free()
        # ---- C2 on the torus in slabs: L = 1024 x 1024 x (1024 N), planes split over the ranks, ghost planes by ncclSend/Recv ----
        from cuda_fortran_mc_simulation_spin_b200 import ising_periodic_gpu_m
        try:
            m = ising_periodic_gpu_m.ising_periodic_gpu().init_distributed(1024, 1024, 1024 * world, KBT, SEED)
            nall = m.nall()
            m.update_n(3); m.sync()
            blocks = T.run(lambda: m.update_n(K))
            e = entry(f"Ising 3D Metropolis, int8, TRUE PERIODIC 1024x1024x{1024 * world} in {world} slab(s) of planes, kbt={KBT}, all-up start",
                      nall, blocks, K, BYTES_PER_FLIP, "step", {"sites_per_gpu": nall // world})

            def loopt():
                s = 0
                for _ in range(K):
                    m.update()
                    s += m.calc_magne_sum() + m.calc_energy_sum()
                return s
            m.update(); m.measure()
            b2 = T.run(loopt)
            e["e2e"] = {"value": nall * K / (statistics.median(b2) * 1e6), "unit": UNIT, "ms_per_step": statistics.median(b2) / K,
                        "note": "update + calc_magne_sum + calc_energy_sum per MCS (sums all-reduced with NCCL every MCS)"}
            out["C2_ising3d_periodic_slabs"] = e
            del m
        except Exception as ex:   # a secondary configuration must not take the headline line down with it
            out["C2_ising3d_periodic_slabs"] = {"error": f"{type(ex).__name__}: {ex}"}
        _free()
        return out

    # ---- C2': the headline lattice on the bit-packed (multi-spin coded) storage, one bit per site (BASELINE.md C2') ----
    from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m
    BPF_BITS = 3.0 / 8.0
    for key, mk, label, kern in (
            ("C2p_ising3d_bitpacked_1023x1023x1024", lambda: ising3d_gpu_m.ising3d_gpu().init_packed(NX, NY, NZ, KBT, SEED),
             f"Ising 3D Metropolis, ONE BIT per site (multi-spin coded), helical {NX}x{NY}x{NZ}, kbt={KBT}, all-up start", "bits_pass_kernel<6>"),
            ("C5p_ising2d_bitpacked_65537x65536", lambda: ising2d_gpu_m.ising2d_gpu().init_packed(65537, 65536, KBT2, SEED),
             f"Ising 2D Metropolis, ONE BIT per site, helical 65537x65536, kbt={KBT2}, all-up start", "bits_pass_kernel<4>")):
        m = mk()
        nall = m.nall()
        kp = K if key.startswith("C2p") else max(2, K // 4)
        m.update_n(3); m.sync()
        m.set_timing(True)
        blocks = T.run(lambda: m.update_n(kp))
        n_pass, pass_ms = m.get_timing()
        m.set_timing(False)
        e = entry(label, nall, blocks, kp, BPF_BITS, "step, against the bit-packed layout's own 3/8 B per flip (the pass is ALU-bound: bit-serial accept test)",
                  {"roofline_kernel": _kernel_roof(nall / 2, BPF_BITS, n_pass, pass_ms, peak, kern),
                   "roofline_vs_int8_bytes": _roof(nall * kp / (statistics.median(blocks) * 1e6), BYTES_PER_FLIP, peak,
                                                   "the same flips/ns expressed in the int8 layout's 3 B per flip (> 1 means faster than an int8 kernel at its HBM roofline could be)")})

        def loopp():
            s = 0
            for _ in range(kp):
                m.update()
                s += m.calc_magne_sum() + m.calc_energy_sum()
            return s
        b2 = T.run(loopp)
        e["e2e"] = {"value": nall * kp / (statistics.median(b2) * 1e6), "unit": UNIT, "ms_per_step": statistics.median(b2) / kp,
                    "note": "update + calc_magne_sum + calc_energy_sum per MCS through the module API (separate measurement kernel)"}
        out[key] = e
        del m
        _free()

    # ---- C2 / C5 on the torus: L = 1024^3 itself and 65536^2 with true periodic boundaries ("1024^3 periodic",
    #      "65536^2 periodic"; not a reference module -- its helical types need odd nx) ----
    from cuda_fortran_mc_simulation_spin_b200 import ising_periodic_gpu_m
    for key, dims, kbt, label, kern in (
            ("C2_ising3d_periodic_1024x1024x1024", (1024, 1024, 1024), KBT,
             f"Ising 3D Metropolis, int8, TRUE PERIODIC 1024x1024x1024 (the north star's L = 1024^3), kbt={KBT}, all-up start", "torus_strip_kernel<6, METROPOLIS, 8 rows, R = 32>"),
            ("C5_ising2d_periodic_65536x65536", (65536, 65536, 0), KBT2,
             f"Ising 2D Metropolis, int8, TRUE PERIODIC 65536x65536, kbt={KBT2}, all-up start", "torus_strip_kernel<4, METROPOLIS, 8 rows>")):
        m = ising_periodic_gpu_m.ising_periodic_gpu().init(*dims, kbt, SEED)
        nall = m.nall()
        kp = K if dims[2] else max(2, K // 4)
        m.update_n(3); m.sync()
        m.set_timing(True)
        blocks = T.run(lambda: m.update_n(kp))
        n_pass, pass_ms = m.get_timing()
        m.set_timing(False)
        e = entry(label, nall, blocks, kp, BYTES_PER_FLIP, "step (2 colour-pass launches per step, nothing else on the stream: no halo)",
                  {"roofline_kernel": _kernel_roof(nall / 2, BYTES_PER_FLIP, n_pass, pass_ms, peak, kern)})

        def loopt():
            s = 0
            for _ in range(kp):
                m.update()
                s += m.calc_magne_sum() + m.calc_energy_sum()
            return s
        m.update(); m.measure()
        b2 = T.run(loopt)
        e["e2e"] = {"value": nall * kp / (statistics.median(b2) * 1e6), "unit": UNIT, "ms_per_step": statistics.median(b2) / kp,
                    "note": "update + calc_magne_sum + calc_energy_sum per MCS through the module API (E and M accumulated by the second colour pass, "
                            "stored to pinned host memory by the kernel)"}
        out[key] = e
        del m
        _free()

    # ---- C1: Ising 2D at the reference driver's defaults (app/ising2d_gpu_relaxation.f90:6-12) ----
    m = ising2d_gpu_m.ising2d_gpu().init(1001, 1000, KBT2, SEED)
    k1 = max(K, 200)
    m.update_n(50); m.sync()
    blocks = T.run(lambda: m.update_n(k1), min_seconds=0.3)
    e = entry("Ising 2D Metropolis, helical 1001x1000 (reference default), one sample, update_n: whole sweeps in one cooperative launch",
              m.nall(), blocks, k1, BYTES_PER_FLIP, "step (launch/barrier-latency bound at this size, not HBM)")

    def loop1():
        s = 0
        for _ in range(k1):
            m.update()
            s += m.calc_magne_sum() + m.calc_energy_sum()
        return s
    b2 = T.run(loop1, min_seconds=0.3)
    e["e2e"] = {"value": m.nall() * k1 / (statistics.median(b2) * 1e6), "unit": UNIT, "us_per_mcs": statistics.median(b2) / k1 * 1e3,
                "note": "update + calc_magne_sum + calc_energy_sum per MCS through the module API"}
    out["C1_ising2d_1001x1000"] = e
    del m
    mb = ising2d_gpu_m.ising2d_gpu().init_multi(1001, 1000, KBT2, SEED, 128)
    mb.update_n(5); mb.sync()
    blocks = T.run(lambda: mb.update_n(K), min_seconds=0.3)
    out["C1_ising2d_1001x1000_batch128"] = entry("the same lattice, 128 independent samples per launch (the drivers' tot_sample loop, batched)",
                                                 mb.nall() * 128, blocks, K, BYTES_PER_FLIP, "step")
    del mb
    _free()

    # ---- C3: XY periodic 16384^2, fp32 angles, from disorder (app/xy2d_periodic_gpu_over_relaxation.f90) ----
    x = xy2d_periodic_gpu_m.xy2d_gpu().init(16384, 16384, 0.89, SEED)
    x.set_random_spin()
    k3 = max(2, K // 2)
    x.update_n(3); x.sync()
    n = x.nall()
    blocks = T.run(lambda: x.update_n(k3))
    out["C3_xy_16384_metropolis"] = entry("XY periodic 16384x16384 fp32 angles, kbt=0.89, from disorder: Metropolis sweeps",
                                          n, blocks, k3, BYTES_PER_FLIP_XY, "step (2 xy_strip_kernel launches per step, nothing else on the stream)")
    blocks = T.run(lambda: x.update_over_relaxation(k3))
    out["C3_xy_16384_over_relaxation"] = entry("the same lattice: over-relaxation steps", n, blocks, k3, BYTES_PER_FLIP_XY,
                                               "step (2 xy_strip_kernel launches per step)")

    def mcs3():
        s = 0.0
        for _ in range(k3):
            x.update(); x.update_over_relaxation(1)
            s += x.calc_energy_sum() + x.calc_magne_sum() + x.calc_magne_y_sum()
        return s
    blocks = T.run(mcs3)
    out["C3_xy_16384_driver_mcs"] = entry("the driver's MCS: 1 Metropolis + 1 over-relaxation + E, Mx, My read back every MCS (2 x nall attempts)",
                                          2 * n, blocks, k3, BYTES_PER_FLIP_XY, "step")
    del x
    _free()

    # ---- C4: q = 6 clock, 16384^2, tableall / dual-lattice semantics, batches of samples; helical clock_gpu_m ----
    for nm in (1, 2, 8):
        c = sixclock(16384, 16384, 0.91, 6, nm, SEED)
        c.update_metropolis_n(2); c.sync()
        k4 = max(2, K // (2 * nm))
        c.set_timing(True)
        blocks = T.run(lambda: c.update_metropolis_n(k4))
        n_l, l_ms = c.get_timing()
        c.set_timing(False)
        out[f"C4_clock6_16384_tableall_batch{nm}"] = entry(
            f"q=6 clock periodic 16384x16384 (clock_tableall / dual-lattice semantics), kbt=0.91, ordered start, {nm} sample(s) per launch",
            c.nall() * nm, blocks, k4, BYTES_PER_FLIP, "step",
            {"roofline_kernel": _kernel_roof(c.nall() * nm / 2, BYTES_PER_FLIP, n_l, l_ms, peak, "sixclock_pass_kernel")})
        del c
        _free()
    h = clock_gpu_m.clock_gpu().init(16385, 16384, 0.91, 6, SEED)
    h.update_n(2); h.sync()
    k4 = max(2, K // 2)
    blocks = T.run(lambda: h.update_n(k4))
    out["C4_clock6_16385x16384_helical"] = entry("q=6 clock helical 16385x16384 (clock_gpu_m), kbt=0.91, ordered start",
                                                 h.nall(), blocks, k4, BYTES_PER_FLIP, "step (2 pass + 2 halo launches per step)")
    del h
    _free()
    return out


def run_ours(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from cuda_fortran_mc_simulation_spin_b200 import _lib, ising3d_gpu_m
    import ctypes as C

    launch_count = _lib.fn("b200mc_launch_count", C.c_ulonglong)
    T = Timer(torch, dist)
    if world > 1:
        # ONE lattice of nz = 1024 N planes, slab-decomposed: every rank owns 1023 x 1023 x 1024 sites
        m = ising3d_gpu_m.ising3d_gpu().init_distributed(NX, NY, NZ * world, KBT, SEED)
    else:
        m = ising3d_gpu_m.ising3d_gpu().init(NX, NY, NZ, KBT, SEED)
    nall = m.nall() // world  # sites per GPU
    K, W = args.steps, max(args.warmup, 3)

    # ---- device-timed sweep throughput (lattice resident in HBM) ----
    m.set_allup_spin()
    m.update_n(W)
    m.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    m.set_timing(True)  # CUDA events around every colour-pass launch, on the launching stream
    l0 = launch_count()
    blocks = T.run(lambda: m.update_n(K))
    l1 = launch_count()
    n_pass, pass_ms = m.get_timing()
    m.set_timing(False)
    ms_med = statistics.median(blocks)
    value = world * nall * K / (ms_med * 1e6)  # flips per ns, whole job

    # ---- end to end through the module API, host-visible results every step ----
    # (the reference drivers' loop: update -> calc_magne_sum -> calc_energy_sum; the API has
    # no per-step host inputs, the two int64 sums are the device->host traffic)
    m.update()
    m.calc_magne_sum()
    acc = [0]

    def e2e_loop():
        for _ in range(K):
            m.update()
            acc[0] += m.calc_magne_sum()
            acc[0] += m.calc_energy_sum()
    eblocks = T.run(e2e_loop)
    e2e_ms = statistics.median(eblocks)
    e2e_value = world * nall * K / (e2e_ms * 1e6)
    p2p = bool(getattr(m, "_p2p", False))
    del m
    _free()
    peak, peak_src = _measured_peak()
    configs = None
    if not args.only_headline:
        configs = bench_other_configs(T, K, peak, world, rank)
    clocks = sampler.stop() if rank == 0 else None  # sampled over all timed regions

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    avg_pass_ms = pass_ms / max(n_pass, 1)
    alg_bytes_per_launch = BYTES_PER_FLIP * (nall / 2)  # one colour pass updates nall/2 sites
    achieved = alg_bytes_per_launch / (avg_pass_ms * 1e-3) / 1e9
    traffic, traffic_src = _ncu_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "ising_pass_kernel<6, METROPOLIS, ORDERED>",
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": avg_pass_ms,
                "launches_timed": n_pass, "kernel_share_of_step": pass_ms / sum(blocks), "peak_source": peak_src,
                "step_frac": value / world * BYTES_PER_FLIP / peak}
    cpu = None
    if world == 1:
        nthr = _host_threads()
        r = _oracle_sweep_rate(nthr, steps=3, warmup=1, min_seconds=10.0)
        r1 = _oracle_sweep_rate(1, steps=2, warmup=1, min_seconds=5.0)
        cpu = {"value": r["flips_per_ns"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": (f"oracle (C restatement of src/ising3d_gpu_m.f90) at {CPU_SHAPE[0]}x{CPU_SHAPE[1]}x{CPU_SHAPE[2]}, "
                          f"{r['steps']} MCS in {r['seconds']:.1f} s: uniforms + 2 colour passes + halos + E + M per MCS"),
               "serial_value": r1["flips_per_ns"], "serial_cores": 1, "serial_steps": r1["steps"]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_med / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"Ising 3D checkerboard Metropolis relaxation, int8 lattice, helical {NX}x{NY}x{NZ} per GPU "
                               f"(reference-valid shape next to 1024^3), kbt={KBT}, all-up start, seed {SEED}",
                   "sites_per_gpu": nall, "rng": "Philox4x32-10 in registers, 32-bit lazy uniforms",
                   "l2": "lattice (2 x 536 MB) is 8x larger than L2; no flush needed",
                   "timing": f"median of {len(blocks)} blocks of exactly {K} steps (min {min(blocks) / K:.4f}, max {max(blocks) / K:.4f} ms/step), "
                             "each bracketed by barrier + synchronize, CUDA events, max over ranks",
                   "parallelism": "1 GPU" if world == 1 else
                   f"one {NX}x{NY}x{NZ * world} lattice in {world} slabs (one process per GPU), halo exchange per colour pass: "
                   + ("boundary results stored straight into the neighbours' halos over NVLink by the colour-pass kernel (CUDA IPC peer memory)"
                      if p2p else "NCCL send/recv on a second stream, overlapped with the interior launch")
                   + ", observables summed across the ranks"},
        "blocks": len(blocks),
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16,
                "steps": K, "blocks": len(eblocks), "ms_per_step": e2e_ms / K,
                "note": "update + calc_magne_sum + calc_energy_sum through the module API every MCS "
                        "(E and M accumulated by the second colour pass, read back and synchronised every step); "
                        "the API takes no host input per step, results are two int64"},
        "gpu_launches": int(l1 - l0),
        "clocks": clocks,
        "checksum": acc[0],
    }
    if configs is not None:
        line["configs"] = configs
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--only-headline", action="store_true", help="skip the other BASELINE configs (C1, C3, C4, C5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
