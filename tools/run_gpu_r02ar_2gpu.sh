#!/bin/bash
# round 2, call ar (2 GPUs): torus p2p slabs with the flag wait in front of the boundary planes of the next pass: parity + timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ising_torus_slab.py -q -rA > gpurun_out/r02ar_torus_slab_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ar_torus_slab_2gpu.log
grep -E "passed|failed|rc=|Error|error|torus slab ok" gpurun_out/r02ar_torus_slab_2gpu.log | tail -8
: > gpurun_out/r02ar_torus_slab_time.log
for w in 0 1; do
  echo "B200MC_TORUS_WAIT_IN_PUSH=$w" >> gpurun_out/r02ar_torus_slab_time.log
  B200MC_TORUS_WAIT_IN_PUSH=$w timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2963$w tools/torus_slab_time.py 2>&1 | grep "torus slabs" >> gpurun_out/r02ar_torus_slab_time.log
done
cat gpurun_out/r02ar_torus_slab_time.log
timeout 200 python tools/quick_torus3.py 2>&1 | tail -1
