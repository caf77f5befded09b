#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_xy.py tests/test_gpu_golden.py tests/test_gpu_curand_stream.py -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
tail -5 gpurun_out/r02o_pytest.log
timeout 300 python tools/quick_models.py > gpurun_out/r02o_quick_models.log 2>&1; head -3 gpurun_out/r02o_quick_models.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02o_bench_ref.json 2>&1; tail -1 gpurun_out/r02o_bench_ref.json | cut -c1-900
