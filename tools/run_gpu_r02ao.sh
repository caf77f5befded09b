#!/bin/bash
# round 2, call ao: 2D torus with the row pitch compiled in (16384^2, 65536^2) vs as a kernel argument
mkdir -p gpurun_out
cat > /tmp/t2d.py <<'PY'
import os, sys, time
sys.path.insert(0, ".")
import torch
from cuda_fortran_mc_simulation_spin_b200 import ising_periodic_gpu_m as M
for dims, n in (((65536, 65536, 0), 12), ((16384, 16384, 0), 100)):
    m = M.ising_periodic_gpu().init(*dims, 2.26918531421, 42)
    m.update_n(5); m.sync(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); m.update_n(n); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n)
    print(os.path.basename(os.environ.get("B200MC_SO") or "default"), dims, f"min {min(ts):.4f} ms/MCS = {m.nall() / min(ts) / 1e6:.0f} flips/ns ({m.nall() / min(ts) / 1e6 * 3 / 6456.5 * 100:.1f} %)", m.measure(), flush=True)
    del m
PY
: > gpurun_out/r02ao_torus2d_pitch.log
for v in "" norc2d; do
  so=""; [ -n "$v" ] && so="$PWD/_ab/libb200mc_$v.so"
  B200MC_SO=$so timeout 300 python /tmp/t2d.py >> gpurun_out/r02ao_torus2d_pitch.log 2>&1
done
cat gpurun_out/r02ao_torus2d_pitch.log
