"""GPU: XY periodic with the rows split over the ranks (slabs along y) against the one-GPU run of the same
lattice.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_xy_slabs_equal_the_one_gpu_run():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
         "--master-port", "29547", os.path.join(ROOT, "tests", "_xy_slab_worker.py")],
        capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert r.stdout.count("xy slab ok") == 3
