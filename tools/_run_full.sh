timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; tail -1 gpurun_out/bench_r01c.json | cut -c1-1500
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
