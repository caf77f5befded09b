timeout 600 python -m pytest tests/test_gpu_clock.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
timeout 120 python tools/clock_time.py 2>&1 | tail -1
B200MC_CLOCK_DIRECT=0 timeout 120 python tools/clock_time.py 2>&1 | tail -1
