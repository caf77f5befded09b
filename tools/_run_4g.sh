timeout 600 python -m pytest tests/test_gpu_slab.py -x -q -k p2p 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/bench_4g_r01c.json 2> gpurun_out/bench_4g_r01c.err; tail -1 gpurun_out/bench_4g_r01c.json | cut -c1-400
