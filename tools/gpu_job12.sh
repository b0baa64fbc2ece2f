#!/bin/bash
# round-2 job 12: bottleneck probes of K1a (no coefficient stores; fewer warps) and K1b at lower occupancy
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
A="--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts --no-verify"
timeout 1500 python tools/sweep_variants.py run --bench-args "$A" nostore a192 a160 a128 b5
