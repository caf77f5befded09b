#!/bin/bash
# round 2, call ag: L2 prefetch distance of the torus strip kernel
mkdir -p gpurun_out
: > gpurun_out/r02ag_torus_variants.log
for v in "" pfd256 pfd512 pfd1024 pfd2048; do
  so=""; [ -n "$v" ] && so="$PWD/_ab/libb200mc_$v.so"
  B200MC_SO=$so timeout 300 python tools/quick_torus3.py >> gpurun_out/r02ag_torus_variants.log 2>&1
done
cat gpurun_out/r02ag_torus_variants.log
