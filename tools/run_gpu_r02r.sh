#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/quick_bits.py > gpurun_out/r02r_bits.log 2>&1; head -3 gpurun_out/r02r_bits.log
B200MC_BITS_SHORT=1 timeout 300 python tools/quick_bits.py > gpurun_out/r02r_bits_short.log 2>&1; head -3 gpurun_out/r02r_bits_short.log
