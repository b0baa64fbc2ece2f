#!/bin/bash
# round-2 job 2: parity of K1a (fixed) + K2 v5 + new entry points, bench line with the new legs, variants
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/j2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j2_pytest.log
tail -8 gpurun_out/j2_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/j2_bench.json 2> gpurun_out/j2_bench.err; echo "bench rc $?"
timeout 1200 python tools/sweep_variants.py run --bench-args "--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-verify" base nolut l8 l12 k2legacy
