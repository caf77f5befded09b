#!/bin/bash
# round 2, call aq (2 GPUs): torus slabs over peer memory: parity with both transports, timing of both
mkdir -p gpurun_out
: > gpurun_out/r02aq_torus_slab_2gpu.log
for tr in p2p nccl; do
  echo "== B200MC_SLAB_TRANSPORT=$tr" >> gpurun_out/r02aq_torus_slab_2gpu.log
  B200MC_SLAB_TRANSPORT=$tr timeout 400 python -m pytest tests/test_gpu_ising_torus_slab.py -q -rA >> gpurun_out/r02aq_torus_slab_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02aq_torus_slab_2gpu.log
done
grep -E "==|passed|failed|rc=|Error|error" gpurun_out/r02aq_torus_slab_2gpu.log | tail -10
: > gpurun_out/r02aq_torus_slab_time.log
for tr in p2p nccl; do
  echo "B200MC_SLAB_TRANSPORT=$tr" >> gpurun_out/r02aq_torus_slab_time.log
  B200MC_SLAB_TRANSPORT=$tr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tools/torus_slab_time.py 2>&1 | grep "torus slabs" >> gpurun_out/r02aq_torus_slab_time.log
done
cat gpurun_out/r02aq_torus_slab_time.log
