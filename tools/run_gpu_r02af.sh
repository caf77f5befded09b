#!/bin/bash
# round 2, call af: compile-time variants of the torus strip kernel
mkdir -p gpurun_out
: > gpurun_out/r02af_torus_variants.log
for v in "" nb2 nb2minb4; do
  so=""; [ -n "$v" ] && so="$PWD/_ab/libb200mc_$v.so"
  B200MC_SO=$so timeout 300 python tools/quick_torus3.py 2d >> gpurun_out/r02af_torus_variants.log 2>&1
done
timeout 600 python -m pytest tests/test_gpu_ising_torus.py -x -q 2>&1 | tail -3; cat gpurun_out/r02af_torus_variants.log
