"""Worker of tests/test_gpu_batch_split.py: a batch of independent clock samples shared out over the ranks of a
torch.distributed group (gloo; the ranks may share one GPU).  No collective on the data path; the gathered
per-sample observables must equal those of one handle holding the whole batch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)) % torch.cuda.device_count())
    dist.init_process_group("gloo")
    from cuda_fortran_mc_simulation_spin_b200 import clock_gpu_multi_m
    from cuda_fortran_mc_simulation_spin_b200._sixclock import sixclock

    n = 5  # 3 + 2 on two ranks
    s = sixclock.distributed(64, 32, 0.91, 6, n, 42)
    s.update_metropolis_n(6)
    e, m = s.calc_energy_all(), s.calc_magne_all()
    c = clock_gpu_multi_m.clock_gpu().init_distributed(33, 32, 0.8, 6, n, 7)
    c.set_random_spin()
    c.update_n(4)
    ce, cm = c.calc_energy_sum_all(), c.calc_magne_sum_all()
    assert len(e) == n and len(ce) == n
    if rank == 0:
        s1 = sixclock(64, 32, 0.91, 6, n, 42)
        s1.update_metropolis_n(6)
        assert np.array_equal(e, s1.calc_energy()) and np.array_equal(m, s1.calc_magne()), (e, s1.calc_energy())
        c1 = clock_gpu_multi_m.clock_gpu().init(33, 32, 0.8, 6, n, 7)
        c1.set_random_spin()
        c1.update_n(4)
        assert np.array_equal(ce, c1.calc_energy_sum()) and np.array_equal(cm, c1.calc_magne_sum()), (ce, c1.calc_energy_sum())
        print("batch split ok: gathered per-sample observables == one-handle batch", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
