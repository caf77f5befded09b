timeout 600 python -m pytest tests/test_gpu_sixclock.py -x -q 2>&1 | tail -4
timeout 120 python tools/ab_models.py 2>&1 | tail -1
B200MC_SIX_DIRECT=0 timeout 600 python -m pytest tests/test_gpu_sixclock.py -x -q 2>&1 | tail -2
B200MC_SIX_DIRECT=0 timeout 120 python tools/ab_models.py 2>&1 | tail -1
