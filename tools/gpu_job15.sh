#!/bin/bash
# round-2 job 15: staging depth 16 / 24, and a source-level ncu capture of the new K1a
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
A="--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts"
timeout 600 python tools/sweep_variants.py run --bench-args "$A" st16 st24
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e-all --no-verify --no-e2e-ts"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_parse_kernel -s 1 -c 1 -o gpurun_out/j15_k1a $B > gpurun_out/j15_a.log 2>&1
ls -la gpurun_out/j15_*
