set -x
timeout 600 python tools/slab_diag.py > gpurun_out/slab_diag1.log 2>&1; tail -20 gpurun_out/slab_diag1.log
