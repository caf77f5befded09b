"""GPU: a batch of independent clock samples split across handles / ranks (the batch dimension of
clock_gpu_multi_m and of the periodic clock modules is embarrassingly parallel -- no exchange) reproduces the
one-handle batch sample for sample."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_offset_reproduces_the_batch():
    from cuda_fortran_mc_simulation_spin_b200 import clock_gpu_multi_m
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock
    whole = sixclock(64, 32, 0.91, 6, 4, 42)
    lo = sixclock(64, 32, 0.91, 6, 2, 42)
    hi = sixclock(64, 32, 0.91, 6, 2, 42).set_sample_offset(2)
    for g in (whole, lo, hi):
        g.update_metropolis_n(5)
    assert np.array_equal(whole.get_sixclock(), np.concatenate([lo.get_sixclock(), hi.get_sixclock()], axis=0))
    assert np.array_equal(whole.calc_energy(), np.concatenate([lo.calc_energy(), hi.calc_energy()]))
    assert not np.array_equal(lo.get_sixclock(), hi.get_sixclock())
    cw = clock_gpu_multi_m.clock_gpu().init(33, 32, 0.8, 6, 3, 7)
    ca = clock_gpu_multi_m.clock_gpu().init(33, 32, 0.8, 6, 1, 7)
    cb = clock_gpu_multi_m.clock_gpu().init(33, 32, 0.8, 6, 2, 7).set_sample_offset(1)
    for g in (cw, ca, cb):
        g.set_random_spin()
        g.update_n(4)
    assert np.array_equal(cw.calc_energy_sum(), np.concatenate([ca.calc_energy_sum(), cb.calc_energy_sum()]))
    assert np.array_equal(cw.calc_magne_sum(), np.concatenate([ca.calc_magne_sum(), cb.calc_magne_sum()]))


def test_batch_split_over_two_ranks():
    """two processes (gloo), one GPU or two: each runs its share of the samples, observables gathered"""
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29541", os.path.join(ROOT, "tests", "_batch_worker.py")],
        capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert "batch split ok" in r.stdout
