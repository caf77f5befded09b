"""Worker of tests/test_gpu_xy_slab.py: one rank of an XY run with the rows split over the ranks (torchrun, one process
per GPU).  Angles must equal the one-GPU run of the same lattice bit for bit (same random numbers, same arithmetic per
site); E, Mx, My are real64 sums taken in a different order: 1e-9 relative."""
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
    from cuda_fortran_mc_simulation_spin_b200 import xy2d_periodic_gpu_m as xm

    for nx, ny_per in ((64, 36), (250, 64), (1024, 512)):
        ny = ny_per * world
        g = xm.xy2d_gpu().init_distributed(nx, ny, 0.89, 11)
        assert g.ny() == ny_per and g.nall() == nx * ny
        n = nx * ny
        assert g.measure() == (-2.0 * n, 1.0 * n, 0.0)
        g.set_random_spin()
        log = [g.measure()]
        for _ in range(3):
            g.update()
            log.append(g.measure())            # separate measurement kernel first, fused sums afterwards
            g.update_over_relaxation(2)
            log.append(g.measure())
        g.metropolis_by_field(0.7, -0.2)
        g.update_n(2)
        log.append(g.measure())
        ang = g.angles_all()
        g.rotate_summation_magne_toward_xaxis()      # the angle comes from the all-reduced sums (equal to ~1e-16 only)
        rot = g.angles_all()
        del g
        if rank == 0:
            g1 = xm.xy2d_gpu().init(nx, ny, 0.89, 11)
            g1.set_random_spin()
            log1 = [g1.measure()]
            for _ in range(3):
                g1.update()
                log1.append(g1.measure())
                g1.update_over_relaxation(2)
                log1.append(g1.measure())
            g1.metropolis_by_field(0.7, -0.2)
            g1.update_n(2)
            log1.append(g1.measure())
            assert np.array_equal(ang, g1.angles()), "angles differ from the one-GPU run"
            g1.rotate_summation_magne_toward_xaxis()
            close = float(np.abs(((rot.astype(np.float64) - g1.angles() + 0.5) % 1.0) - 0.5).max())
            assert close < 3e-7, close
            for (e, mx, my), (e1, mx1, my1) in zip(log, log1):
                assert abs(e - e1) <= 1e-6 * abs(e1) + 1e-6 * n, (e, e1)
                assert abs(mx - mx1) <= 1e-6 * n and abs(my - my1) <= 1e-6 * n, ((mx, my), (mx1, my1))
            print("xy slab ok", (nx, ny), f"angles identical; after the rotation max difference {close:.1e} turns", flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
