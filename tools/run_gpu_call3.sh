#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
timeout 120 python tools/curand_probe.py > gpurun_out/r02_curand_probe.log 2>&1; cat gpurun_out/r02_curand_probe.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_1gpu.json 2> gpurun_out/r02b_bench_1gpu.err; echo "bench rc=$?"
tail -1 gpurun_out/r02b_bench_1gpu.json | cut -c1-6000
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02b_bench_ref.json 2>&1; tail -1 gpurun_out/r02b_bench_ref.json | cut -c1-1500
# ncu: launch list of the bench command (headline only, short), then full captures of the two default clock kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_bench_launches.csv python bench.py --steps 5 --warmup 3 --only-headline > gpurun_out/r02b_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:clock_pass -s 2 -c 2 -o gpurun_out/prof_r02b_clock python tools/prof_models.py clock > gpurun_out/r02b_ncu_clock.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sixclock_pass -s 2 -c 2 -o gpurun_out/prof_r02b_sixclock python tools/prof_models.py sixclock > gpurun_out/r02b_ncu_sixclock.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
