#!/bin/bash
# round 2, call x: the whole GPU suite on the self-cleaning / folded-sums build
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/r02x_gputests_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_gputests_1gpu.log
tail -12 gpurun_out/r02x_gputests_1gpu.log
