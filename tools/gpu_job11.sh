#!/bin/bash
# round-2 job 11: K1 with the context in kernel-parameter space, K1a v3 step (dequantisation in K1a) + ring look-ahead; occupancy sweep
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_host_mirror_gpu.py -m gpu -x -q > gpurun_out/j11_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j11_pytest.log
tail -4 gpurun_out/j11_pytest.log
A="--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts"
timeout 1500 python tools/sweep_variants.py run --bench-args "$A" base a256 a288 a320 a192x5 noahead nopf oldp b8 b6x5
