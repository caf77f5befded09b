#!/usr/bin/env python
"""bench.py -- headline benchmark of the checkerboard lattice-spin MC sweep.

Metric (BASELINE.json): attempted spin flips / ns, device-timed, and the
fraction of the HBM roofline.  Workload at N GPUs (weak scaling): BASELINE
config 2, the Ising 3D checkerboard Metropolis relaxation on an int8 lattice,
at the reference-valid helical shape next to 1024^3: 1023 x 1023 x 1024 per
GPU (1024^3 itself is rejected by the reference's linear-index colouring,
SURVEY.md Q1), kbt = 4.51152, all-up start (app/ising3d_gpu_relaxation.f90).
A "step" is one MCS = one `update()` = nall attempted flips per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of
the reference's own algorithm (oracle/, kind "port": the reference is CUDA
Fortran and cannot be built in this image) on the host cores.
"""
from __future__ import annotations

import argparse
import json
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
SEED = 42
METRIC = "spin_flips_per_ns"
UNIT = "flips/ns"
BYTES_PER_FLIP = 3.0  # int8 two-pass checkerboard: read other colour, read own, write own
CPU_SHAPE = (255, 255, 256)  # bounded CPU sample: same model / temperature / start, 1/64 of the sites


def _measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic():
    """DRAM bytes per launch of the pass kernel from the committed ncu capture (or None)."""
    p = os.path.join(ROOT, "profiles", "r01_ising3d_pass_traffic.json")
    try:
        with open(p) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""

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
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _oracle_sweep_rate(steps: int, warmup: int, min_seconds: float = 0.0):
    """Time the CPU restatement of the reference's per-MCS work on all host threads:
    draw nall uniforms (the reference's curandGenerate, src/ising3d_gpu_m.f90:179), two
    colour passes + halo copies (:180-187), then the two reductions the drivers call every
    MCS (calc_magne_sum, calc_energy_sum; app/ising3d_gpu_relaxation.f90:43-45)."""
    import numpy as np

    from oracle import oracle as O

    O.build()
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
    while done < steps or (time.perf_counter() - t0) < min_seconds:
        one()
        done += 1
    dt = time.perf_counter() - t0
    return {"flips_per_ns": m.nall() * done / dt / 1e9, "steps": done, "seconds": dt,
            "cores": O.max_threads(), "nall": m.nall()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = _oracle_sweep_rate(args.steps, args.warmup)
    sample = (f"Ising 3D helical {CPU_SHAPE[0]}x{CPU_SHAPE[1]}x{CPU_SHAPE[2]} (1/64 of the per-GPU lattice), "
              f"{r['steps']} MCS, per MCS: draw nall uniforms + 2 colour passes + halo copies + E and M reductions, "
              f"OpenMP on {r['cores']} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["flips_per_ns"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["seconds"] / r["steps"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32 spins / f64 uniforms",
        "data": "synthetic",
        "config": {"workload": f"Ising 3D checkerboard Metropolis, helical {NX}x{NY}x{NZ} per GPU, kbt={KBT}, all-up start",
                   "note": "CPU port of the reference algorithm timed on a bounded sample; flips/ns is size-independent"},
        "cpu_baseline": {"value": r["flips_per_ns"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["flips_per_ns"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


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
    if world > 1:
        # ONE lattice of nz = 1024 N planes, slab-decomposed: every rank owns 1023 x 1023 x 1024 sites,
        # boundary vectors exchanged after each colour pass (NCCL send/recv over NVLink, overlapped)
        m = ising3d_gpu_m.ising3d_gpu().init_distributed(NX, NY, NZ * world, KBT, SEED)
    else:
        m = ising3d_gpu_m.ising3d_gpu().init(NX, NY, NZ, KBT, SEED)
    nall = m.nall() // world  # sites per GPU
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed sweep throughput (lattice resident in HBM) ----
    m.set_allup_spin()
    m.update_n(W)
    m.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    m.set_timing(True)  # CUDA events around every colour-pass launch, on the launching stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = launch_count()
    barrier()
    e0.record()
    m.update_n(K)
    e1.record()
    barrier()
    l1 = launch_count()
    ms = e0.elapsed_time(e1)
    n_pass, pass_ms = m.get_timing()
    m.set_timing(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * nall * K / (ms_max * 1e6)  # flips per ns, whole job

    # ---- end to end through the module API, host-visible results every step ----
    # (the reference drivers' loop: update -> calc_magne_sum -> calc_energy_sum; the API has
    # no per-step host inputs, the two int64 sums are the device->host traffic)
    Ke = max(K, 5)
    m.update()
    m.calc_magne_sum()
    barrier()
    e0.record()
    acc = 0
    for _ in range(Ke):
        m.update()
        acc += m.calc_magne_sum()
        acc += m.calc_energy_sum()
    e1.record()
    barrier()
    ms_e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * nall * Ke / (float(t.item()) * 1e6)
    clocks = sampler.stop() if rank == 0 else None  # sampled over both timed regions (device-timed sweeps + e2e loop)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = _measured_peak()
    avg_pass_ms = pass_ms / max(n_pass, 1)
    alg_bytes_per_launch = BYTES_PER_FLIP * (nall / 2)  # one colour pass updates nall/2 sites
    achieved = alg_bytes_per_launch / (avg_pass_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": _ncu_traffic(), "kernel": "ising_pass_kernel<6, METROPOLIS>",
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": avg_pass_ms,
                "launches_timed": n_pass, "kernel_share_of_step": pass_ms / ms, "peak_source": peak_src}
    cpu = None
    if world == 1:
        r = _oracle_sweep_rate(steps=3, warmup=1, min_seconds=10.0)
        cpu = {"value": r["flips_per_ns"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": (f"oracle (C restatement of src/ising3d_gpu_m.f90) at {CPU_SHAPE[0]}x{CPU_SHAPE[1]}x{CPU_SHAPE[2]}, "
                          f"{r['steps']} MCS in {r['seconds']:.1f} s: uniforms + 2 colour passes + halos + E + M per MCS")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"Ising 3D checkerboard Metropolis relaxation, int8 lattice, helical {NX}x{NY}x{NZ} per GPU "
                               f"(reference-valid shape next to 1024^3), kbt={KBT}, all-up start, seed {SEED}",
                   "sites_per_gpu": nall, "rng": "Philox4x32-10 in registers, 32-bit lazy uniforms",
                   "l2": "lattice (2 x 536 MB) is 8x larger than L2; no flush needed",
                   "parallelism": "1 GPU" if world == 1 else
                   f"one {NX}x{NY}x{NZ * world} lattice in {world} slabs (one process per GPU), halo exchange per colour pass: "
                   + ("boundary results stored straight into the neighbours' halos over NVLink by the colour-pass kernel (CUDA IPC peer memory)"
                      if getattr(m, "_p2p", False) else "NCCL send/recv on a second stream, overlapped with the interior launch")
                   + ", observables all-reduced (NCCL)"},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16,
                "steps": Ke, "note": "update + calc_magne_sum + calc_energy_sum through the module API every MCS "
                                     "(E and M accumulated by the second colour pass, read back and synchronised every step); "
                                     "the API takes no host input per step, results are two int64"},
        "gpu_launches": int(l1 - l0),
        "clocks": clocks,
        "checksum": acc,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
