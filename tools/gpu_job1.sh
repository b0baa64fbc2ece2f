#!/bin/bash
# round-2 job 1: parity of the new K1a on the GPU, then the knob sweep
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/j1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j1_pytest.log
tail -5 gpurun_out/j1_pytest.log
timeout 900 python tools/sweep_variants.py run base h24 h16 h12
timeout 900 python tools/sweep_variants.py run --check h8 l11h16
timeout 600 python tools/sweep_variants.py run nopfh16 t256h16
