#!/bin/bash
# round 2, call h: bit-packed Ising -- parity, timing, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ising_bits.py tests/test_c_consumer.py tests/test_gpu_ising.py -q -x > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -30 gpurun_out/r02h_pytest.log
timeout 300 python tools/quick_bits.py > gpurun_out/r02h_quick_bits.log 2>&1; cat gpurun_out/r02h_quick_bits.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bits_pass -s 6 -c 2 -o gpurun_out/prof_r02h_bits python tools/quick_bits.py > gpurun_out/r02h_ncu_bits.log 2>&1
ls -la gpurun_out/*r02h*
