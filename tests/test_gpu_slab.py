"""GPU parity of the slab-decomposed (multi-GPU) Ising path: torchrun, one process per GPU,
bit-exact against the CPU oracle of the global lattice.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_slab_two_ranks_bit_exact(transport):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
         "--master-port", "29533", os.path.join(ROOT, "tests", "_slab_worker.py")],
        capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, B200MC_SLAB_TRANSPORT=transport))
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0
    assert "N-rank run == 1-GPU run" in r.stdout
    assert r.stdout.count("slab ok: large slabs == 1-GPU run") == 6
    assert f"interleaved run == 1-GPU run, transport {transport} p2p active: {transport == 'p2p'}" in r.stdout
