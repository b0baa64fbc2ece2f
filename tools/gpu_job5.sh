#!/bin/bash
# round-2 job 5: suite on the restructured TS / prefix kernels + tcgen05 experiment; K2 variants; bench with the TS leg; ncu of the experiment + K2
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/j5_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j5_pytest.log
grep -E "level-1|tcgen05|passed|failed|rc |Error" gpurun_out/j5_pytest.log | tail -10
timeout 900 python tools/sweep_variants.py run --check --bench-args "--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-verify --no-e2e-ts" base
timeout 900 python tools/sweep_variants.py run --bench-args "--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-verify --no-e2e-ts" k2a k2t704 k2a704
timeout 600 python -m pytest tests/test_composite_gpu.py -m gpu -q > gpurun_out/j5_k2a_check.log 2>&1 <<< "" ; EF_LIB=libespflix_b200.k2a.so timeout 600 python -m pytest tests/test_composite_gpu.py -m gpu -q >> gpurun_out/j5_k2a_check.log 2>&1; tail -2 gpurun_out/j5_k2a_check.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/j5_bench.json 2> gpurun_out/j5_bench.err; echo "bench rc $?"; tail -2 gpurun_out/j5_bench.err
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e-all --no-verify --no-e2e-ts"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_composite_kernel -s 2 -c 1 -o gpurun_out/j5_k2n $B > gpurun_out/j5_ncu_k2n.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_idct_tc_kernel -c 1 -o gpurun_out/j5_idct python -m pytest tests/test_idct_tc_gpu.py -m gpu -q -k sparse > gpurun_out/j5_ncu_idct.log 2>&1
ls -la gpurun_out/j5_* gpurun_out/idct_tc_*
