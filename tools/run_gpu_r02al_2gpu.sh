#!/bin/bash
# round 2, call al (2 GPUs): torus slabs after the priority-stream / compile-time-pitch change: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ising_torus_slab.py -q -rA > gpurun_out/r02al_torus_slab_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02al_torus_slab_2gpu.log
grep -E "torus slab ok|passed|failed|rc=" gpurun_out/r02al_torus_slab_2gpu.log | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/torus_slab_time.py 2>&1 | grep "torus slabs" | tee gpurun_out/r02al_torus_slab_time.log
