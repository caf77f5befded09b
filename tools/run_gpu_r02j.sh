#!/bin/bash
# round 2, call j: whole suite + smoke + full bench line + bits timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -16 gpurun_out/r02j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02j_smoke.log; tail -2 gpurun_out/r02j_smoke.log
timeout 300 python tools/quick_bits.py > gpurun_out/r02j_quick_bits.log 2>&1; cat gpurun_out/r02j_quick_bits.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02j_bench_1gpu.json 2> gpurun_out/r02j_bench_1gpu.err; echo "bench rc=$?"
tail -1 gpurun_out/r02j_bench_1gpu.json | cut -c1-1500
