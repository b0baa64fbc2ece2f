#!/bin/bash
# round-2 job 22: K2 v8 (persistent CTAs with ONE band each: no CTA turnover, occupancy kept) against v6; ncu of the tcgen05 experiment
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 600 python -m pytest tests/test_composite_gpu.py -m gpu -x -q > gpurun_out/j22_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j22_pytest.log
EF_LIB=libespflix_b200.k2p5.so timeout 600 python -m pytest tests/test_composite_gpu.py -m gpu -x -q >> gpurun_out/j22_pytest.log 2>&1; echo "pytest k2p5 rc $?" >> gpurun_out/j22_pytest.log
grep -E "passed|failed|rc " gpurun_out/j22_pytest.log
A="--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts --no-verify"
timeout 900 python tools/sweep_variants.py run --bench-args "$A" base k2p5 k2p4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_idct_tc_kernel -c 1 -o gpurun_out/j22_idct python -m pytest tests/test_idct_tc_gpu.py -m gpu -q -k sparse > gpurun_out/j22_ncu_idct.log 2>&1
ls -la gpurun_out/j22_*
