#!/bin/bash
# round 2, call at (4 GPUs): torus slab parity over peer memory with distinct neighbours
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ising_torus_slab.py -q -rA > gpurun_out/r02at_torus_slab_4gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02at_torus_slab_4gpu.log
grep -E "torus slab ok|passed|failed|rc=|Error|error" gpurun_out/r02at_torus_slab_4gpu.log | tail -8
