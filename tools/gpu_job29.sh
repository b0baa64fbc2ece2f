#!/bin/bash
# round-2 job 29: last unroll sweep of K1a (steps per vote x flush sources per pass) on the final code
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
A="--steps 20 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts"
timeout 900 python tools/sweep_variants.py run --bench-args "$A" base u4 fu8 u4fu8 u2fu8 base
