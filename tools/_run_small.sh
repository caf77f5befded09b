timeout 120 python tools/small_lattice.py 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_ising.py tests/test_gpu_golden.py -x -q 2>&1 | tail -6
