#!/bin/bash
# round 2, call k: XY contract v2 + bits fused measure: parity, timings, sanitizer, helical clock block-size A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
tail -12 gpurun_out/r02k_pytest.log
timeout 300 python tools/quick_models.py > gpurun_out/r02k_quick_models.log 2>&1; cat gpurun_out/r02k_quick_models.log
B200MC_CLOCK_THREADS=1024 timeout 300 python tools/clock_time.py > gpurun_out/r02k_clock_1024.log 2>&1; cat gpurun_out/r02k_clock_1024.log | tail -4
timeout 300 python tools/quick_bits.py > gpurun_out/r02k_quick_bits.log 2>&1; head -3 gpurun_out/r02k_quick_bits.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r02k_san_memcheck.log 2>&1; tail -4 gpurun_out/r02k_san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r02k_san_racecheck.log 2>&1; tail -4 gpurun_out/r02k_san_racecheck.log
