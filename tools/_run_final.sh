timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 120 python tools/ab_models.py 2>&1 | tail -1
timeout 120 python tools/quick_models.py 2>&1 | grep -i "xy" | tail -6
timeout 300 python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; tail -1 gpurun_out/bench_r01d.json | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench_r01d.log 2>&1; tail -2 gpurun_out/ncu_bench_r01d.log | cut -c1-200
