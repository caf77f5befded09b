#!/bin/bash
# round 2, call d: clock contract v2 kernels -- parity tests first, then timings, then the whole suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sixclock.py tests/test_gpu_clock.py tests/test_gpu_golden.py tests/test_gpu_batch_split.py -q -x > gpurun_out/r02d_pytest_clock.log 2>&1; echo "pytest clock rc=$?" >> gpurun_out/r02d_pytest_clock.log
tail -25 gpurun_out/r02d_pytest_clock.log
timeout 300 python tools/quick_sixclock.py > gpurun_out/r02d_quick_sixclock.log 2>&1; cat gpurun_out/r02d_quick_sixclock.log
timeout 300 python tools/quick_models.py > gpurun_out/r02d_quick_models.log 2>&1; cat gpurun_out/r02d_quick_models.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -8 gpurun_out/r02d_pytest.log
