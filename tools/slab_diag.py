"""Slab-pass variants, timed.  One GPU: plain vs the self-neighbour experiment (B200MC_TUNE bit 4).
Under torchrun: real slabs.  Prints ms/sweep and checks E, M against the plain 1-GPU run where results are valid."""
import os, sys
sys.path.insert(0, ".")
import torch

def timeit(m, n):
    m.update_n(3); m.sync()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.update_n(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m, ising2d_gpu_m
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    variants = [("tune", "nb")]
    if world == 1:
        variants = [(0, None), (16, None), (16, 16), (16, 32), (16, 48), (16, 148), (16, 222), (16 | 1024, 148), (16 | 512, None)]
    else:
        variants = [(0, None), (0, 48), (0, 120), (0, 148), (1024, 148), (512, None), (4, None), (512 | 4, None)]
    ref = None
    for tune, nb in variants:
        os.environ["B200MC_TUNE"] = str(tune)
        if nb is None: os.environ.pop("B200MC_SLAB_NB", None)
        else: os.environ["B200MC_SLAB_NB"] = str(nb)
        for kind in ("3d", "2d"):
            if kind == "3d":
                args = (1023, 1023, 1024 * world, 4.51152, 42)
                m = ising3d_gpu_m.ising3d_gpu()
            else:
                args = (65537, 65536 * world, 2.26918531421, 42)
                m = ising2d_gpu_m.ising2d_gpu()
            m = m.init_distributed(*args) if world > 1 else m.init(*args)
            ms = timeit(m, 40 if kind == "3d" else 12)
            em = m.measure()
            if rank == 0:
                print(f"world={world} {kind} TUNE={tune} NB={nb}: {ms:.4f} ms/sweep {m.nall()/ms/1e6:.1f} flips/ns  E,M={em}", flush=True)
            del m
            if dist: dist.barrier()

main()
