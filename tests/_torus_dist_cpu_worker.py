"""Worker of tests/test_dist_cpu.py::test_torus_slab_protocol_two_gloo_ranks: one of WORLD_SIZE gloo ranks on the CPU.
A numpy model of the torus slab pass as csrc/ising_torus.cu runs it -- every rank owns nz / P planes plus one ghost plane below
and above, updates its planes with the GLOBAL plane index in the colour parity and in the uniform it draws, then sends its
first / last owned plane of the colour just updated to rank - 1 / rank + 1 (ring of ranks) -- against the oracle of the
global lattice: merged spins bit-exact after every sweep, the all-reduced {X, sum s} give the oracle's E and M."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from oracle import oracle as O

    if rank == 0:
        O.build()
    dist.barrier()
    prev, nxt = (rank - 1) % world, (rank + 1) % world

    def exchange(loc):
        """loc: [nzl + 2][ny][nx] with ghost planes 0 and nzl + 1; first owned plane -> prev's upper ghost, last -> next's lower"""
        first, last = torch.from_numpy(loc[1].copy()), torch.from_numpy(loc[-2].copy())
        low, high = torch.empty_like(first), torch.empty_like(first)
        ops = [dist.P2POp(dist.isend, first, prev, tag=1), dist.P2POp(dist.isend, last, nxt, tag=2),
               dist.P2POp(dist.irecv, high, nxt, tag=1), dist.P2POp(dist.irecv, low, prev, tag=2)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        loc[0], loc[-1] = low.numpy(), high.numpy()

    for (nx, ny, nzl, method) in [(32, 4, 2, 0), (64, 6, 3, 0), (32, 8, 5, 1)]:
        nz = nzl * world
        kbt = 4.51152
        o = O.ising_periodic_gpu().init(nx, ny, nz, kbt, 42)
        o.set_random_spin()
        glob = o.spins().reshape(nz, ny, nx)
        z0 = rank * nzl
        loc = np.empty((nzl + 2, ny, nx), dtype=np.int32)
        loc[1:-1] = glob[z0:z0 + nzl]
        exchange(loc)
        assert np.array_equal(loc[0], glob[(z0 - 1) % nz]) and np.array_equal(loc[-1], glob[(z0 + nzl) % nz])
        zz, yy, xx = np.meshgrid(np.arange(z0, z0 + nzl), np.arange(ny), np.arange(nx), indexing="ij")   # GLOBAL plane index
        w = o.w.reshape(2, 7) if method == 0 else None
        for sweep in range(3):
            u = O.isingp_uniforms(42, o.draw_, nx, ny, nz).reshape(nz, ny, nx)[z0:z0 + nzl]   # the rank's share of the global stream
            (o.update_heatbath if method else o.update)()
            for colour in (0, 1):
                own = loc[1:-1]
                nsum = (np.roll(own, 1, axis=2) + np.roll(own, -1, axis=2) + np.roll(own, 1, axis=1) + np.roll(own, -1, axis=1)
                        + loc[:-2] + loc[2:])                      # z neighbours: the planes below / above, ghosts at the ends
                mask = ((xx + yy + zz) & 1) == colour
                if method == 0:
                    acc = u <= w[own, nsum]
                    loc[1:-1] = np.where(mask & acc, 1 - own, own)
                else:
                    loc[1:-1] = np.where(mask, (u <= o.pup[nsum]).astype(np.int32), own)
                exchange(loc)                                      # (both colours live in one array here: one exchange serves both)
            merged = [torch.empty((nzl, ny, nx), dtype=torch.int32) for _ in range(world)]
            dist.all_gather(merged, torch.from_numpy(loc[1:-1].copy()))
            assert np.array_equal(torch.cat(merged).numpy().ravel(), o.spins()), (nx, ny, nzl, method, sweep)
            # observables: X = unequal neighbour pairs seen from the colour-1 sites this rank owns, sum s over its sites
            own = loc[1:-1]
            nb = [np.roll(own, 1, axis=2), np.roll(own, -1, axis=2), np.roll(own, 1, axis=1), np.roll(own, -1, axis=1), loc[:-2], loc[2:]]
            c1 = ((xx + yy + zz) & 1) == 1
            X = sum(int(((n != own) & c1).sum()) for n in nb)
            t = torch.tensor([X, int(own.sum())], dtype=torch.int64)
            dist.all_reduce(t)
            N = nx * ny * nz
            assert (-3 * N + 2 * int(t[0]), 2 * int(t[1]) - N) == o.measure(), (nx, ny, nzl, method, sweep)
    dist.barrier()
    if rank == 0:
        print(f"torus dist cpu ok world={world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
