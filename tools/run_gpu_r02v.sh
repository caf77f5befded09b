#!/bin/bash
# round 2, call v: self-cleaning launches + fused pass with batched loads vs the memset/copy form
mkdir -p gpurun_out
timeout 900 python tools/ab_ising.py base,SELF_CLEAN=0 > gpurun_out/r02v_ab.log 2>&1
grep -v "^$" gpurun_out/r02v_ab.log | tail -40
