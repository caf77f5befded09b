#!/bin/bash
# round 2, call as (2 GPUs): torus p2p slabs, push kernel on 296 blocks: parity + timing; single-GPU torus number of the same build
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ising_torus_slab.py -q -rA > gpurun_out/r02as_torus_slab_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02as_torus_slab_2gpu.log
grep -E "passed|failed|rc=|Error|error" gpurun_out/r02as_torus_slab_2gpu.log | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tools/torus_slab_time.py 2>&1 | grep "torus slabs" | tee gpurun_out/r02as_torus_slab_time.log
timeout 200 python tools/quick_torus3.py 2>&1 | tail -1 | tee -a gpurun_out/r02as_torus_slab_time.log
