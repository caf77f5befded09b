#!/bin/bash
# round 2, call ak (2 GPUs): torus slabs -- parity against the oracle of the global lattice, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ising_torus_slab.py -q -rA > gpurun_out/r02ak_torus_slab_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ak_torus_slab_2gpu.log
grep -E "torus slab ok|passed|failed|Error|error|assert|rc=" gpurun_out/r02ak_torus_slab_2gpu.log | tail -14
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02ak_bench_2gpu.log 2>&1
echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02ak_bench_2gpu.log'):
    if l.startswith('{"metric"'):
        open('gpurun_out/r02ak_bench_2gpu.json','w').write(l)
        d=json.loads(l); print("value", d['value'], "e2e", d['e2e']['value'])
        for k,v in d['configs'].items(): print(k, round(v['value'],1), round(v['ms_per_step'],4), round(v.get('e2e',{}).get('value',0),1))
PY
tail -5 gpurun_out/r02ak_bench_2gpu.log | cut -c1-300
