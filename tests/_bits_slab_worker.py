"""Worker of tests/test_gpu_bits_slab.py: one rank of a slab-decomposed bit-packed Ising run (launched by torchrun, one
process per GPU).  Every rank checks the merged configuration and the all-reduced observables against the CPU oracle of
the GLOBAL lattice run on the bit-packed contract's uniforms, bit for bit, after every sweep."""
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
    for kind, shape, start in [("3d", (15, 17, 256 * world), "allup"), ("3d", (31, 31, 512 * world), "random"),
                               ("2d", (33, 256 * world), "random"), ("2d", (255, 512 * world), "allup")]:
        if kind == "3d":
            g = ising3d_gpu_m.ising3d_gpu().init_packed_distributed(*shape, KBT3, 42)
            o = O.ising3d_gpu().init(*shape, KBT3, 42)
        else:
            g = ising2d_gpu_m.ising2d_gpu().init_packed_distributed(*shape, KBT2, 42)
            o = O.ising2d_gpu().init(*shape, KBT2, 42)
        assert g.rank_info() == (rank, world)
        n, draw = g.nall(), 0
        if start == "random":
            g.set_random_spin(); o.set_random_spin(O.isingbits_uniforms(42, draw, n, init=True)); draw += 1
        assert np.array_equal(g.spins(), o.spins()), (kind, shape, "initial")
        for sweep in range(4):
            g.update(); o.update(randoms=O.isingbits_uniforms(42, draw, n)); draw += 1
            assert np.array_equal(g.spins(), o.spins()), (kind, shape, start, sweep)
            assert g.measure() == (o.calc_energy_sum(), o.calc_magne_sum()), (kind, shape, start, sweep)
        s = o.spins()
        g.set_allup_spin(); g.set_spins(s)
        assert np.array_equal(g.spins(), s) and g.measure() == (o.calc_energy_sum(), o.calc_magne_sum())
        del g
        if rank == 0:
            print("bits slab ok", kind, shape, start, flush=True)
    # an N-rank run equals the 1-GPU run of the same lattice
    g = ising3d_gpu_m.ising3d_gpu().init_packed_distributed(63, 65, 256 * world, KBT3, 7)
    g.set_random_spin(); g.update_n(5)
    em = g.measure()
    if rank == 0:
        g1 = ising3d_gpu_m.ising3d_gpu().init_packed(63, 65, 256 * world, KBT3, 7)
        g1.set_random_spin(); g1.update_n(5)
        assert g1.measure() == em
        print("bits slab ok: N-rank run == 1-GPU run", em, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
