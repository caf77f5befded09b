#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02p_slab_2gpu.log 2>&1
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_xy_slab.py tests/test_gpu_batch_split.py -q -rA >> gpurun_out/r02p_slab_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02p_slab_2gpu.log
grep -E "xy slab ok|passed|failed|PASS|FAIL|rc=" gpurun_out/r02p_slab_2gpu.log | tail -12
