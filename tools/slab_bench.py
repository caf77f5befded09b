"""torchrun worker: time update_n of the slab-decomposed Ising 3D headline lattice (per-rank 1023x1023x1024)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
from cuda_fortran_mc_simulation_spin_b200 import ising3d_gpu_m, ising2d_gpu_m
kind = sys.argv[1] if len(sys.argv) > 1 else "3d"
if kind == "3d":
    g = ising3d_gpu_m.ising3d_gpu().init_distributed(1023, 1023, 1024 * world, 4.51152, 42)
else:
    g = ising2d_gpu_m.ising2d_gpu().init_distributed(65537, 65536 * world, 2.26918531421, 42)
n = 30
g.update_n(5); g.sync(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.update_n(n); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
import ctypes as C
from cuda_fortran_mc_simulation_spin_b200 import _lib
w = _lib.fn("b200mc_debug_slab_wait_ns", C.c_ulonglong, C.c_void_p)(g._h)
tw = torch.tensor([float(w)], device="cuda"); dist.all_reduce(tw, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"  max flag-wait per pass {tw.item() / (2 * (n + 5)) / 1e3:.1f} us;", end=" ")
    print(f"{kind} world={world} TUNE={os.environ.get('B200MC_TUNE','0')} ILEAVE={os.environ.get('B200MC_ILEAVE','1')} "
          f"TRANSPORT={os.environ.get('B200MC_SLAB_TRANSPORT','p2p')}: {t.item():.4f} ms/MCS  {g.nall()/t.item()/1e6:.1f} flips/ns", flush=True)
dist.barrier(); dist.destroy_process_group()
