#!/bin/bash
# round 2, call f: clock contract v3 kernels -- parity, timings, ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sixclock.py tests/test_gpu_clock.py tests/test_gpu_golden.py tests/test_gpu_batch_split.py tests/test_gpu_curand_stream.py -q > gpurun_out/r02f_pytest_clock.log 2>&1; echo "pytest clock rc=$?" >> gpurun_out/r02f_pytest_clock.log
tail -12 gpurun_out/r02f_pytest_clock.log
timeout 300 python tools/quick_sixclock.py > gpurun_out/r02f_quick_sixclock.log 2>&1; cat gpurun_out/r02f_quick_sixclock.log
B200MC_SIX_THREADS=1024 timeout 300 python tools/quick_sixclock.py > gpurun_out/r02f_quick_sixclock_1024.log 2>&1; cat gpurun_out/r02f_quick_sixclock_1024.log
timeout 300 python tools/quick_models.py > gpurun_out/r02f_quick_models.log 2>&1; cat gpurun_out/r02f_quick_models.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:clock_pass -s 2 -c 2 -o gpurun_out/prof_r02f_clock python tools/prof_models.py clock > gpurun_out/r02f_ncu_clock.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sixclock_pass -s 2 -c 2 -o gpurun_out/prof_r02f_sixclock python tools/prof_models.py sixclock > gpurun_out/r02f_ncu_sixclock.log 2>&1
ls -la gpurun_out/*r02f*
