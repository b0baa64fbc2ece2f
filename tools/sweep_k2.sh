#!/bin/bash
# tools/sweep_k2.sh — K2 variant check (run on the GPU box via gpurun): chroma LUT through L1 (default) vs in shared memory
for defs in "" "-DEF_K2_LUT_SMEM"; do
  EF_NVCC_DEFS="$defs" python -c "
from espflix_b200 import build; build.build_cuda(force=True)"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['composite']; print('K2 [$defs]', round(c['ntsc']['ms_per_launch'],3), round(c['ntsc']['frac'],3), round(c['pal']['ms_per_launch'],3), round(c['pal']['frac'],3), 'K1', round(d['roofline']['k1_ms_per_step'],2))"
done
python -c "
from espflix_b200 import build; build.build_cuda(force=True)"
