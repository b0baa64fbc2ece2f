#!/bin/bash
# round-2 job 7: suite on the new K0 scan / smem quantiser table / 16x2 clamp / poster scroll; K1 knob sweep; tcgen05 experiment at throughput size
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/j7_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j7_pytest.log
grep -E "level-1|tcgen05|passed|failed|rc |Error" gpurun_out/j7_pytest.log | tail -10
timeout 1500 python tools/sweep_variants.py run --bench-args "--steps 10 --warmup 3 --no-cpu --no-e2e-all --no-verify --no-e2e-ts" base nopf t256 t256nopf pin32
