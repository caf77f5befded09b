"""Worker of tests/test_gpu_ising_torus_slab.py: one rank of a slab-decomposed periodic (torus) Ising 3D run (launched by
torchrun, one process per GPU).  Every rank checks the merged configuration and the all-reduced observables against the CPU
oracle of the GLOBAL lattice, bit for bit, after every sweep; then an N-rank run against the 1-GPU run of the same lattice."""
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
    from cuda_fortran_mc_simulation_spin_b200 import ising_periodic_gpu_m as M
    from oracle import oracle as O

    if rank == 0:
        O.build()
    dist.barrier()
    KBT = 4.51152
    # planes per rank: 2 (boundary launch only), 3 (odd: the row parity needs the global plane index), 8; 8- and 2-row tickets; two strips
    for shape, method, start in [((1024, 8, 2 * world), 0, "random"), ((1024, 6, 3 * world), 0, "allup"),
                                 ((1024, 16, 8 * world), 1, "random"), ((2048, 8, 4 * world), 0, "random")]:
        g = M.ising_periodic_gpu().init_distributed(*shape, KBT, 42)
        o = O.ising_periodic_gpu().init(*shape, KBT, 42)
        g.set_method(method)
        step = o.update_heatbath if method else o.update
        r, n, z0, nzl = g.rank_info()
        assert (r, n) == (rank, world) and nzl == shape[2] // world and z0 == rank * nzl
        assert g.measure() == o.measure()
        if start == "random":
            g.set_random_spin(); o.set_random_spin()
        assert np.array_equal(g.spins(), o.spins()), (shape, "initial")
        for sweep in range(5):
            g.update(); step()
            assert g.measure() == o.measure(), (shape, method, start, sweep)      # sweep 0: measure kernel; then fused
            assert np.array_equal(g.spins(), o.spins()), (shape, method, start, sweep)
        g.update_n(3); step(); step(); step()
        assert np.array_equal(g.spins(), o.spins()) and g.measure() == o.measure()
        s = o.spins()
        g.set_allup_spin(); g.set_spins(s)
        assert np.array_equal(g.spins(), s) and g.measure() == o.measure()
        g.update(); step()
        assert np.array_equal(g.spins(), o.spins()) and g.measure() == o.measure()
        del g
        if rank == 0:
            print("torus slab ok", shape, method, start, flush=True)
    g = M.ising_periodic_gpu().init_distributed(1024, 64, 16 * world, KBT, 7)
    g.set_random_spin(); g.update_n(5)
    em = g.measure()
    if rank == 0:
        g1 = M.ising_periodic_gpu().init(1024, 64, 16 * world, KBT, 7)
        g1.set_random_spin(); g1.update_n(5)
        assert g1.measure() == em
        print("torus slab ok: N-rank run == 1-GPU run", em, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
