#!/bin/bash
# round 2, call ah: L2 prefetch distance, torus (more points) and helical
mkdir -p gpurun_out
: > gpurun_out/r02ah_variants.log
for v in pfd768 pfd1536 r4pfd2048; do
  B200MC_SO=$PWD/_ab/libb200mc_$v.so timeout 300 python tools/quick_torus3.py >> gpurun_out/r02ah_variants.log 2>&1
done
for v in "" hpfd16384 hpfd65536 hpfd131072; do
  so=""; [ -n "$v" ] && so="$PWD/_ab/libb200mc_$v.so"
  B200MC_SO=$so timeout 300 python tools/quick_hel3.py >> gpurun_out/r02ah_variants.log 2>&1
done
cat gpurun_out/r02ah_variants.log
