#!/bin/bash
# round-2 job 27: K1a staging row addressed from the ring address, per-macroblock table row pointer: parity + bench
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_host_mirror_gpu.py -m gpu -x -q > gpurun_out/j27_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j27_pytest.log
tail -3 gpurun_out/j27_pytest.log
A="--steps 20 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts"
timeout 900 python tools/sweep_variants.py run --bench-args "$A" base
