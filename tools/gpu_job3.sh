#!/bin/bash
# round-2 job 3: audio parity on the GPU + ncu captures of the current K1a / K1b / K2 (analysis)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_audio_gpu.py tests/test_composite_gpu.py -m gpu -q > gpurun_out/j3_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j3_pytest.log
tail -4 gpurun_out/j3_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e-all --no-verify"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_parse_kernel -s 1 -c 1 -o gpurun_out/j3_k1a $B > gpurun_out/j3_ncu_k1a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_recon_kernel -s 13 -c 2 -o gpurun_out/j3_k1b $B > gpurun_out/j3_ncu_k1b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_composite_kernel -s 2 -c 1 -o gpurun_out/j3_k2n $B > gpurun_out/j3_ncu_k2n.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_scan_kernel -s 1 -c 1 -o gpurun_out/j3_k0 $B > gpurun_out/j3_ncu_k0.log 2>&1
ls -la gpurun_out/j3_*
