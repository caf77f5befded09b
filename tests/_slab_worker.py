"""Worker of tests/test_gpu_slab.py: one rank of a slab-decomposed Ising run (launched by torchrun,
one process per GPU).  Every rank checks the merged configuration and the all-reduced observables
against the CPU oracle of the GLOBAL lattice, bit for bit, after every sweep."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from cuda_fortran_mc_simulation_spin_b200 import ising2d_gpu_m, ising3d_gpu_m
    from oracle import oracle as O

    if rank == 0:
        O.build()
    dist.barrier()
    KBT3, KBT2 = 4.51152, 2.26918531421
    cases = [("3d", (31, 31, 64 * world), "allup", 0), ("3d", (31, 31, 128 * world), "random", 0),
             ("3d", (15, 17, 256 * world), "random", 1), ("2d", (33, 32 * world), "random", 0),
             ("2d", (33, 512 * world), "allup", 0), ("2d", (129, 256 * world), "random", 1)]
    for kind, shape, start, method in cases:
        if kind == "3d":
            g = ising3d_gpu_m.ising3d_gpu().init_distributed(*shape, KBT3, 42)
            o = O.ising3d_gpu().init(*shape, KBT3, 42)
        else:
            g = ising2d_gpu_m.ising2d_gpu().init_distributed(*shape, KBT2, 42)
            o = O.ising2d_gpu().init(*shape, KBT2, 42)
        assert g.rank_info() == (rank, world)
        if method:
            g.set_method(1)
        if start == "random":
            g.set_random_spin(); o.set_random_spin()
        assert np.array_equal(g.spins(), o.spins()), (kind, shape, "initial")
        for sweep in range(4):
            g.update()
            o.update_heatbath() if method else o.update()
            assert np.array_equal(g.spins(), o.spins()), (kind, shape, start, sweep)
            assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum()), (kind, shape, start, sweep)
        # explicit uniforms (reference-stream mode) and set_spins through the slab path
        rng = np.random.default_rng(5)
        u = 1.0 - rng.random(g.nall())
        if not method:
            g.update_with_randoms(u); o.update(randoms=u)
            assert np.array_equal(g.spins(), o.spins()), (kind, shape, "randoms")
        s = o.spins()
        g.set_allup_spin()
        g.set_spins(s)
        assert np.array_equal(g.spins(), s), (kind, shape, "set_spins")
        assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())
        del g
        if rank == 0:
            print("slab ok", kind, shape, start, "heatbath" if method else "metropolis", flush=True)
    # an N-rank run equals the 1-GPU run of the same lattice (the random stream does not depend on the ranks)
    g = ising3d_gpu_m.ising3d_gpu().init_distributed(63, 65, 64 * world, KBT3, 7)
    g.update_n(5)
    em = g.measure()
    if rank == 0:
        g1 = ising3d_gpu_m.ising3d_gpu().init(63, 65, 64 * world, KBT3, 7)
        g1.update_n(5)
        assert g1.measure() == em
        print("slab ok: N-rank run == 1-GPU run", em, flush=True)
    # longer asynchronous run through the overlapped path (boundary launch + interior launch per colour):
    # sweeps, re-initialisation and measurements interleaved, against the 1-GPU run of the same lattice
    g = ising3d_gpu_m.ising3d_gpu().init_distributed(127, 127, 96 * world, KBT3, 11)
    log = []
    for rep in range(3):
        g.update_n(20)
        log.append(g.measure())
        g.set_random_spin()
        g.update_n(7)
        log.append(g.measure())
        g.set_allup_spin()
    sp = g.spins()
    if rank == 0:
        g1 = ising3d_gpu_m.ising3d_gpu().init(127, 127, 96 * world, KBT3, 11)
        log1 = []
        for rep in range(3):
            g1.update_n(20)
            log1.append(g1.measure())
            g1.set_random_spin()
            g1.update_n(7)
            log1.append(g1.measure())
            g1.set_allup_spin()
        assert log == log1, (log, log1)
        assert np.array_equal(sp, g1.spins())
        print("slab ok: interleaved run == 1-GPU run, transport", os.environ.get("B200MC_SLAB_TRANSPORT", "p2p"),
              "p2p active:", getattr(g, "_p2p", False), flush=True)
    # slabs large enough for the shared-ticket pass (boundary + interior blocks of two launches resident together, one
    # ticket counter): configuration and per-sweep fused E, M against the 1-GPU run of the same lattice, both methods
    # (255 x 263: H = 33533 vectors, H % 128 = 125 -- the halo ends 48 bytes into a 128-byte line that also holds owned
    # vectors, the layout in which a boundary block could hit a stale L1 line if it read the halo through the nc path)
    for kind, shape, kbt in (("3d", (255, 255, 256 * world), KBT3), ("3d", (255, 263, 256 * world), KBT3), ("2d", (4097, 6144 * world), KBT2)):
        mod = ising3d_gpu_m.ising3d_gpu if kind == "3d" else ising2d_gpu_m.ising2d_gpu
        for method in (0, 1):
            g = mod().init_distributed(*shape, kbt, 5)
            g.set_method(method)
            g.set_random_spin()
            log = []
            for sweep in range(6):
                g.update()
                log.append(g.measure())
            g.update_n(5)
            sp = g.spins()
            del g
            if rank == 0:
                g1 = mod().init(*shape, kbt, 5)
                g1.set_method(method)
                g1.set_random_spin()
                log1 = []
                for sweep in range(6):
                    g1.update()
                    log1.append(g1.measure())
                g1.update_n(5)
                assert log == log1, (kind, method, log, log1)
                assert np.array_equal(sp, g1.spins()), (kind, method)
                del g1
                print("slab ok: large slabs == 1-GPU run", kind, shape, "heatbath" if method else "metropolis", flush=True)
            dist.barrier()
    # the device-side driver loop through the slab path: one all-reduce of the whole series
    g = ising3d_gpu_m.ising3d_gpu().init_distributed(31, 31, 64 * world, KBT3, 3)
    e, m = g.run_relaxation(5)
    o = O.ising3d_gpu().init(31, 31, 64 * world, KBT3, 3)
    for i in range(5):
        o.update()
        assert (int(e[i]), int(m[i])) == (o.calc_energy_sum(), o.calc_magne_sum()), i
    assert np.array_equal(g.spins(), o.spins())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
