"""CPU, world_size 2 over gloo: host-side logic of the N > 1 path (slab geometry, ownership, the spins()
merge protocol, the partition of the all-reduced observable sums, rendezvous plumbing, loud failure
without a device).  The GPU twin is tests/test_gpu_slab.py."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_host_logic_two_gloo_ranks():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29547", os.path.join(ROOT, "tests", "_dist_cpu_worker.py")],
        capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert "dist cpu ok world=2" in r.stdout


def test_torus_slab_protocol_two_gloo_ranks():
    """the slab protocol of the periodic (torus) Ising module -- ghost planes, global plane index in parity and RNG, the
    partition of the observable sums -- as a numpy model on two gloo ranks against the oracle of the global lattice;
    the GPU twin is tests/test_gpu_ising_torus_slab.py"""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29549", os.path.join(ROOT, "tests", "_torus_dist_cpu_worker.py")],
        capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert "torus dist cpu ok world=2" in r.stdout


def test_split_samples_covers_the_batch_once():
    """host logic of the batch split (clock samples across ranks): contiguous, disjoint, complete, balanced"""
    from cuda_fortran_mc_simulation_spin_b200.clock_gpu_multi_m import split_samples
    for n in (1, 2, 5, 8, 13):
        for world in (1, 2, 3, 8):
            parts = [split_samples(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
