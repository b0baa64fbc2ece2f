#!/bin/bash
# round-2 job 14: K1a with 16-byte bitstream chunks (cp.async.cg) and shared-memory staged, warp-coalesced list output
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_host_mirror_gpu.py -m gpu -x -q > gpurun_out/j14_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j14_pytest.log
tail -4 gpurun_out/j14_pytest.log
A="--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts"
timeout 1500 python tools/sweep_variants.py run --bench-args "$A" base st0 es4 st64 st48 oldst c3
