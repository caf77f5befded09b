#!/bin/bash
# round 2, call au: last check of the committed build: smoke() + the torus parity tests (without the full-size case)
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_ising_torus.py tests/test_gpu_golden.py -q -k "not full_size" 2>&1 | tail -2
