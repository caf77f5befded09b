#!/bin/bash
mkdir -p gpurun_out
free -g | head -2
timeout 1200 python -m pytest tests/test_gpu_ising_bits.py tests/test_gpu_sixclock.py tests/test_gpu_clock.py -q --durations=6 -k "headline or c4 or large_q" > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
tail -14 gpurun_out/r02n_pytest.log
