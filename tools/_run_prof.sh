set -x
N="ncu --set full --clock-control none --import-source on -f"
timeout 300 $N -k regex:xy_strip -s 4 -c 4 -o gpurun_out/prof_r1i_xy python tools/prof_models.py xy > gpurun_out/ncu_i_xy.log 2>&1
timeout 300 $N -k regex:sixclock_pass -s 2 -c 2 -o gpurun_out/prof_r1i_sixclock python tools/prof_models.py sixclock > gpurun_out/ncu_i_six.log 2>&1
B200MC_TUNE=16 timeout 300 $N -k regex:"ising_slab_kernel|ising_pass_kernel" -s 4 -c 4 -o gpurun_out/prof_r1i_slabself python tools/prof_models.py ising_slabself > gpurun_out/ncu_i_slab.log 2>&1
timeout 300 $N -k regex:ising_coop -s 1 -c 1 -o gpurun_out/prof_r1i_coop python tools/prof_models.py ising_small > gpurun_out/ncu_i_coop.log 2>&1
tail -2 gpurun_out/ncu_i_*.log
