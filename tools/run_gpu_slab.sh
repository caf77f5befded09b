#!/bin/bash
# multi-GPU parity + scaling: gpurun --gpus N --timeout 1800 -- 'bash tools/run_gpu_slab.sh N'
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02_slab_${N}gpu.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_slab.py tests/test_gpu_xy_slab.py tests/test_gpu_batch_split.py -q -rA >> gpurun_out/r02_slab_${N}gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_slab_${N}gpu.log
grep -E "slab ok|passed|failed|PASS|FAIL|rc=" gpurun_out/r02_slab_${N}gpu.log | tail -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.log 2>&1
echo "bench rc=$?" >> gpurun_out/r02_bench_${N}gpu.log
tail -2 gpurun_out/r02_bench_${N}gpu.log
