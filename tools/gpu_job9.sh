#!/bin/bash
# round-2 job 9: parity of K1a v3 (raw tokens, dequantisation in K1b), then the header-batch sweep on both parsers
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_host_mirror_gpu.py -m gpu -x -q > gpurun_out/j9_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j9_pytest.log
tail -4 gpurun_out/j9_pytest.log
A="--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts"
timeout 1500 python tools/sweep_variants.py run --bench-args "$A" old base v3h24 v3h16 v3h12 v3h8 oldh16
