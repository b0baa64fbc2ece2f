#!/bin/bash
# round-2 job 17: K1a steps per vote (2 is the new default) 3 / 4, occupancy again now that the kernel is issue bound
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
A="--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts"
timeout 1500 python tools/sweep_variants.py run --bench-args "$A" base u3 u4 u2t256 u2t192
