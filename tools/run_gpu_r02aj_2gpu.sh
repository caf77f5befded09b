#!/bin/bash
# round 2, call aj (2 GPUs): slab parity of every multi-GPU path + bench on the end-of-session library
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02aj_slab_parity_2gpu.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_slab.py tests/test_gpu_xy_slab.py tests/test_gpu_batch_split.py tests/test_gpu_bits_slab.py -q -rA >> gpurun_out/r02aj_slab_parity_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02aj_slab_parity_2gpu.log
grep -E "passed|failed|rc=" gpurun_out/r02aj_slab_parity_2gpu.log | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02aj_bench_2gpu.log 2>&1
echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02aj_bench_2gpu.log'):
    if l.startswith('{"metric"'):
        open('gpurun_out/r02aj_bench_2gpu.json','w').write(l)
        d=json.loads(l); print("value", d['value'], "e2e", d['e2e']['value'])
        for k,v in d['configs'].items(): print(k, round(v['value'],1), round(v['ms_per_step'],4))
PY
