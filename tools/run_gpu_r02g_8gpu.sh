#!/bin/bash
# round 2, 8 GPUs: slab parity on 8 ranks (both transports), XY slabs, then bench.py --gpus 8 (headline + C5)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02g_slab_${N}gpu.log 2>&1
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_xy_slab.py -q -rA >> gpurun_out/r02g_slab_${N}gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02g_slab_${N}gpu.log
grep -E "slab ok|passed|failed|PASS|FAIL|rc=" gpurun_out/r02g_slab_${N}gpu.log | tail -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02g_bench_${N}gpu.log 2>&1
echo "bench rc=$?" >> gpurun_out/r02g_bench_${N}gpu.log
tail -2 gpurun_out/r02g_bench_${N}gpu.log | cut -c1-3000
