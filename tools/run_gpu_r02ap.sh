#!/bin/bash
# round 2, call ap: end-of-session validation on one B200: whole GPU suite, smoke, bench line, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/r02ap_gputests_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ap_gputests_1gpu.log
tail -4 gpurun_out/r02ap_gputests_1gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02ap_bench_1gpu.json 2> gpurun_out/r02ap_bench_1gpu.err ) 2>&1 | grep real
tail -1 gpurun_out/r02ap_bench_1gpu.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02ap_bench_launches.csv python bench.py --steps 5 --warmup 3 --only-headline > gpurun_out/r02ap_ncu_bench.log 2>&1; echo "ncu rc=$?"
