#!/bin/bash
# round 2, call ac: ncu --set full of the batched torus strip kernel (plain and fused-measurement)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:torus_strip -s 42 -c 2 -o gpurun_out/prof_r02ac_torus3d python tools/prof_torus.py torus3d > gpurun_out/r02ac_ncu_torus3d.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:torus_strip -s 44 -c 4 -o gpurun_out/prof_r02ac_torus3d_fused python tools/prof_torus.py torus3d fused > gpurun_out/r02ac_ncu_torus3d_fused.log 2>&1
ls -la gpurun_out/*r02ac*.ncu-rep
