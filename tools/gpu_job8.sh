#!/bin/bash
# round-2 job 8: ncu captures (source-level) of the current K1a / K1b (P picture) / K0 / K2 + the launch list of one bench step
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e-all --no-verify --no-e2e-ts"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_parse_kernel -s 1 -c 1 -o gpurun_out/j8_k1a $B > gpurun_out/j8_ncu_k1a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_recon_kernel -s 13 -c 2 -o gpurun_out/j8_k1b $B > gpurun_out/j8_ncu_k1b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j8_launches.csv $B > gpurun_out/j8_launches.log 2>&1
ls -la gpurun_out/j8_*
