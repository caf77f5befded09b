timeout 600 python -m pytest tests/test_gpu_xy.py tests/test_gpu_xyh.py -x -q 2>&1 | tail -5
timeout 120 python tools/ab_models.py 2>&1 | tail -1
