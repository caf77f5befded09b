#!/bin/bash
# round 2, call i: bit-packed Ising with the scalar tail -- parity, timing A/B of the occupancy variant, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ising_bits.py tests/test_c_consumer.py -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -8 gpurun_out/r02i_pytest.log
timeout 300 python tools/quick_bits.py > gpurun_out/r02i_quick_bits.log 2>&1; cat gpurun_out/r02i_quick_bits.log
B200MC_BITS_MINB=4 timeout 300 python tools/quick_bits.py > gpurun_out/r02i_quick_bits_minb4.log 2>&1; head -2 gpurun_out/r02i_quick_bits_minb4.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bits_pass -s 6 -c 2 -o gpurun_out/prof_r02i_bits python tools/quick_bits.py > gpurun_out/r02i_ncu_bits.log 2>&1
