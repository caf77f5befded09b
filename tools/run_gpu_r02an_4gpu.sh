#!/bin/bash
# round 2, call an (4 GPUs): the torus slab parity worker with distinct neighbours + the bench line at N = 4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ising_torus_slab.py -q -rA > gpurun_out/r02an_torus_slab_4gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02an_torus_slab_4gpu.log
grep -E "torus slab ok|passed|failed|rc=" gpurun_out/r02an_torus_slab_4gpu.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02an_bench_4gpu.log 2>&1
echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02an_bench_4gpu.log'):
    if l.startswith('{"metric"'):
        open('gpurun_out/r02an_bench_4gpu.json','w').write(l)
        d=json.loads(l); print("value", d['value'], "e2e", d['e2e']['value'])
        for k,v in d['configs'].items(): print(k, v.get('error') or (round(v['value'],1), round(v['ms_per_step'],4), round(v.get('e2e',{}).get('value',0),1)))
PY
