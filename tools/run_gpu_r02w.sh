#!/bin/bash
# round 2, call w: compile-time variants of the ticket pass (queue look per ticket, 256-vector tickets)
mkdir -p gpurun_out
: > gpurun_out/r02w_ab.log
for v in "" tqt ch256 ch256tqt; do
  so=""; [ -n "$v" ] && so="$PWD/_ab/libb200mc_$v.so"
  echo "== build ${v:-default}" >> gpurun_out/r02w_ab.log
  B200MC_SO=$so timeout 600 python tools/ab_ising.py base >> gpurun_out/r02w_ab.log 2>&1
done
grep -v "^$" gpurun_out/r02w_ab.log | tail -40
