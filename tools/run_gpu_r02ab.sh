#!/bin/bash
# round 2, call ab: periodic (torus) Ising -- parity tests, then timing next to the helical module
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ising_torus.py -x -q --durations=5 > gpurun_out/r02ab_torus_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ab_torus_tests.log
tail -15 gpurun_out/r02ab_torus_tests.log
timeout 600 python tools/quick_torus.py > gpurun_out/r02ab_torus_timing.log 2>&1
cat gpurun_out/r02ab_torus_timing.log | tail -12
