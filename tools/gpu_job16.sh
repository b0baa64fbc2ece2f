#!/bin/bash
# round-2 job 16: K1a (now issue bound): branch-free table index, generic table pointer, 2 steps per vote, header batching again
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 900 python -m pytest tests/test_decode_gpu.py -m gpu -x -q > gpurun_out/j16_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j16_pytest.log
tail -3 gpurun_out/j16_pytest.log
A="--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts"
timeout 1500 python tools/sweep_variants.py run --bench-args "$A" prev base u2 h24 h16 u2h16
