#!/bin/bash
# One gpurun call: the whole -m gpu suite (no -x, so every failure is seen), smoke(), a short bench.
# Usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/run_gpu_suite.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/suite_env.log 2>&1
free -g >> gpurun_out/suite_env.log; nproc >> gpurun_out/suite_env.log
timeout 1800 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/suite_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/suite_pytest.log
tail -40 gpurun_out/suite_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/suite_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/suite_smoke.log
tail -3 gpurun_out/suite_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/suite_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/suite_bench.log
tail -3 gpurun_out/suite_bench.log
