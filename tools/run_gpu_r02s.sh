#!/bin/bash
# round 2, call s: A/B of the software-pipelined ticket kernel (B200MC_PIPE) + self-cleaning ticket counters
mkdir -p gpurun_out
timeout 900 python tools/ab_pipe.py 0,2,3,5,6,2m,0 > gpurun_out/r02s_ab_pipe.log 2>&1
cat gpurun_out/r02s_ab_pipe.log | grep -v "^$" | tail -40
