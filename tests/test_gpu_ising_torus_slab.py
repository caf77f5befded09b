"""GPU parity of the slab-decomposed periodic (torus) Ising path: torchrun, one process per GPU, bit-exact against the CPU
oracle of the global lattice.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_torus_slabs_bit_exact():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
         "--master-port", "29547", os.path.join(ROOT, "tests", "_torus_slab_worker.py")],
        capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert r.stdout.count("torus slab ok") == 5 and "N-rank run == 1-GPU run" in r.stdout
