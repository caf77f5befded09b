"""torchrun worker: torus slabs 1024 x 1024 x (1024 N): ms per MCS (update_n) and the drivers' loop, max over ranks"""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
from cuda_fortran_mc_simulation_spin_b200 import ising_periodic_gpu_m as M
m = M.ising_periodic_gpu().init_distributed(1024, 1024, 1024 * world, 4.51152, 42)
m.update_n(5); m.sync()
def block(fn):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t)
n = 40
ts = [block(lambda: m.update_n(n)) / n for _ in range(5)]
def loop():
    for _ in range(n):
        m.update(); m.measure()
m.update(); m.measure()
tl = [block(loop) / n for _ in range(3)]
if rank == 0:
    nall = m.nall()
    print(f"torus slabs x{world}: update_n min {min(ts):.4f} ms/MCS = {nall / min(ts) / 1e6:.0f} flips/ns; loop {min(tl):.4f} ms = {nall / min(tl) / 1e6:.0f} flips/ns; E,M={m.measure()}", flush=True)
dist.destroy_process_group()
