#!/bin/bash
# round-2 job 21: K1a staged-list flush as a fixed walk over the 32 source lanes (immediates instead of ffs), unroll sweep
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 900 python -m pytest tests/test_decode_gpu.py -m gpu -x -q > gpurun_out/j21_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j21_pytest.log
tail -3 gpurun_out/j21_pytest.log
A="--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts"
timeout 900 python tools/sweep_variants.py run --bench-args "$A" base fu1 fu4 fu8 fu4u2
