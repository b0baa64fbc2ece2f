#!/bin/bash
# round-2 job 10: why does header batching lose? source-level ncu captures of K1a v3 with HDR_BATCH 32 and 16 (+ K1b v3)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e-all --no-verify --no-e2e-ts"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_parse_kernel -s 1 -c 1 -o gpurun_out/j10_k1a_v3 $B > gpurun_out/j10_a.log 2>&1
EF_LIB=libespflix_b200.v3h16.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_parse_kernel -s 1 -c 1 -o gpurun_out/j10_k1a_v3h16 $B > gpurun_out/j10_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_recon_kernel -s 13 -c 1 -o gpurun_out/j10_k1b_v3 $B > gpurun_out/j10_c.log 2>&1
ls -la gpurun_out/j10_*
