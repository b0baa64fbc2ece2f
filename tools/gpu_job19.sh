#!/bin/bash
# round-2 job 19: K2 v7 (persistent CTAs, double-buffered bands): parity, then v7 / v6 / 2 CTAs per SM
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 900 python -m pytest tests/test_composite_gpu.py tests/test_host_mirror_gpu.py -m gpu -x -q > gpurun_out/j19_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j19_pytest.log
tail -4 gpurun_out/j19_pytest.log
A="--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts"
timeout 900 python tools/sweep_variants.py run --bench-args "$A" base k2v6 k2c2
