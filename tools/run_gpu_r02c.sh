#!/bin/bash
# round 2, call c: the whole -m gpu suite (no -x), the SFU bias probe, quick model timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -25 gpurun_out/r02c_pytest.log
timeout 60 tools/probes/sfu_bias > gpurun_out/r02c_sfu_bias.log 2>&1; cat gpurun_out/r02c_sfu_bias.log
timeout 300 python tools/quick_models.py > gpurun_out/r02c_quick_models.log 2>&1; cat gpurun_out/r02c_quick_models.log
