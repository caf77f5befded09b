#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_xy.py tests/test_gpu_golden.py tests/test_gpu_xy_slab.py -q -x > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
tail -5 gpurun_out/r02l_pytest.log
timeout 300 python tools/quick_models.py > gpurun_out/r02l_quick_models.log 2>&1; head -3 gpurun_out/r02l_quick_models.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xy_strip -s 4 -c 2 -o gpurun_out/prof_r02l_xy python tools/prof_models.py xy > gpurun_out/r02l_ncu_xy.log 2>&1
