#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clock.py tests/test_gpu_sixclock.py tests/test_gpu_xy.py -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
tail -6 gpurun_out/r02m_pytest.log
timeout 300 python tools/quick_sixclock.py > gpurun_out/r02m_six.log 2>&1; head -2 gpurun_out/r02m_six.log
B200MC_SIX_PREFETCH=1 timeout 300 python tools/quick_sixclock.py > gpurun_out/r02m_six_pf.log 2>&1; head -2 gpurun_out/r02m_six_pf.log
