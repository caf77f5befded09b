"""Worker of tests/test_dist_cpu.py: one of WORLD_SIZE gloo ranks on the CPU (no GPU).  Exercises the
host-side logic of the slab decomposition: slab geometry through the C ABI, site ownership,
the INT32_MIN / elementwise-max merge protocol of spins(), the partition of the observable sums that the
ranks all-reduce, the rendezvous plumbing of init_distributed -- and that the product fails loudly on
every rank when there is no CUDA device."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def owner_of_sites(n_sites, geoms):
    """rank owning each linear site index i (0-based): colour-site k = i >> 1 sits at position p = k % L
    of the folded ring; rank r owns positions [p0_r, p0_r + Lloc_r) of every lane"""
    L = geoms[0]["L"]
    p = (np.arange(n_sites) >> 1) % L
    own = np.full(n_sites, -1, dtype=np.int64)
    for r, g in enumerate(geoms):
        own[(p >= g["p0"]) & (p < g["p0"] + g["Lloc"])] = r
    return own


def main():
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from cuda_fortran_mc_simulation_spin_b200 import B200MCError, ising3d_gpu_m
    from cuda_fortran_mc_simulation_spin_b200._ising_base import slab_geometry, unique_id
    from oracle import oracle as O

    if rank == 0:
        O.build()
    dist.barrier()

    for (nx, ny, nz) in [(31, 31, 64), (15, 17, 256), (33, 64, 0), (129, 256, 0)]:
        # 1. every rank asks the library for ITS slab; together the slabs tile the fold exactly once
        mine = slab_geometry(nx, ny, nz, rank, world)
        geoms = [None] * world
        dist.all_gather_object(geoms, mine)
        L, H = mine["L"], mine["H"]
        assert all(g["L"] == L and g["H"] == H and g["Nc"] == mine["Nc"] for g in geoms)
        assert geoms[0]["p0"] == 0 and sum(g["Lloc"] for g in geoms) == L
        for a, b in zip(geoms, geoms[1:]):
            assert a["p0"] + a["Lloc"] == b["p0"]
        assert all(g["Lloc"] >= H for g in geoms)            # a halo block comes from ONE neighbour
        assert max(g["Lloc"] for g in geoms) - min(g["Lloc"] for g in geoms) <= 1

        # 2. the merge protocol of spins(): non-owned sites read INT32_MIN, elementwise max over ranks
        is3d = nz > 0
        o = (O.ising3d_gpu().init(nx, ny, nz, 4.51152, 42) if is3d else O.ising2d_gpu().init(nx, ny, 2.26918531421, 42))
        o.set_random_spin()
        for _ in range(2):
            o.update()
        full = o.spins()
        halo = nx * ny if is3d else nx
        n = o.nall()
        own = owner_of_sites(n, geoms)
        assert (own >= 0).all()
        local = np.full_like(full, np.iinfo(np.int32).min)
        sel = np.nonzero(own == rank)[0]
        local[halo + sel] = full[halo + sel]
        # halo cells of the reference layout mirror interior sites: they belong to the owner of that site
        lo_src, hi_src = np.arange(n - halo, n), np.arange(0, halo)
        local[np.arange(halo)[own[lo_src] == rank]] = full[np.arange(halo)[own[lo_src] == rank]]
        local[(halo + n + np.arange(halo))[own[hi_src] == rank]] = full[(halo + n + np.arange(halo))[own[hi_src] == rank]]
        t = torch.from_numpy(local.copy())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert np.array_equal(t.numpy(), full)

        # 3. observables: every rank sums over the colour-1 sites it owns (X = unequal neighbours; every bond
        #    has exactly one colour-1 end) and over all its sites (sum s); the all-reduced sums give E and M
        s = full[halo:halo + n].astype(np.int64)
        s01 = s if is3d else (s + 1) // 2
        offs = (1, nx, nx * ny) if is3d else (1, nx)
        i1 = sel[(sel & 1) == 1]
        X = 0
        for d in offs:
            X += int((s01[i1] != s01[(i1 + d) % n]).sum()) + int((s01[i1] != s01[(i1 - d) % n]).sum())
        part = torch.tensor([X, int(s01[sel].sum())], dtype=torch.int64)
        dist.all_reduce(part)
        nnb = 2 * len(offs)
        assert -(nnb // 2) * n + 2 * int(part[0]) == o.calc_energy_sum()
        assert 2 * int(part[1]) - n == o.calc_magne_sum()

    # 4. rendezvous plumbing: rank 0 draws the communicator id, everyone receives the same 128 bytes
    box = [unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ids = [None] * world
    dist.all_gather_object(ids, box[0])
    assert all(i == ids[0] and len(i) == 128 for i in ids)

    # 5. no CPU fallback: without a CUDA device init_distributed fails loudly on EVERY rank (and does not hang)
    if not torch.cuda.is_available():
        try:
            ising3d_gpu_m.ising3d_gpu().init_distributed(31, 31, 64 * world, 4.51152, 42)
            raise SystemExit("init_distributed succeeded without a GPU")
        except B200MCError as e:
            assert "no CPU fallback" in str(e)
    dist.barrier()
    if rank == 0:
        print(f"dist cpu ok world={world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
