#!/bin/bash
# round 2, call am (2 GPUs): torus slabs, one launch + exchange vs boundary split: parity of the default + timing of both
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ising_torus_slab.py -q -rA > gpurun_out/r02am_torus_slab_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02am_torus_slab_2gpu.log
grep -E "passed|failed|rc=" gpurun_out/r02am_torus_slab_2gpu.log | tail -3
: > gpurun_out/r02am_torus_slab_time.log
for sp in 0 1; do
  echo "B200MC_TORUS_SPLIT=$sp" >> gpurun_out/r02am_torus_slab_time.log
  B200MC_TORUS_SPLIT=$sp timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$sp tools/torus_slab_time.py 2>&1 | grep "torus slabs" >> gpurun_out/r02am_torus_slab_time.log
done
cat gpurun_out/r02am_torus_slab_time.log
