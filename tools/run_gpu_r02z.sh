#!/bin/bash
# round 2, call z: ncu --set full of the torus strip kernel and of the helical pass in the same box
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:torus_strip -s 42 -c 2 -o gpurun_out/prof_r02z_torus3d python tools/prof_torus.py torus3d > gpurun_out/r02z_ncu_torus3d.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ising_pass_kernel -s 42 -c 2 -o gpurun_out/prof_r02z_helical3d python tools/prof_torus.py helical3d > gpurun_out/r02z_ncu_helical3d.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r02z_ncu_torus3d.log
