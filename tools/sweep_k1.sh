#!/bin/bash
# tools/sweep_k1.sh — tuning sweep of K1's warps-per-CTA / list-capacity (run on the GPU box via gpurun)
for cfg in "15 80" "15 64" "16 72" "14 96" "16 64"; do
  set -- $cfg
  EF_NVCC_DEFS="-DEF_K1_WARPS=$1 -DEF_K1_LIST=$2" python -c "
from espflix_b200 import build; s=build.build_cuda(force=True, verbose_ptxas=True)
import re; print('cfg $1 $2', re.findall(r'ef_decode_kernel.*?Used (\d+) registers', s, re.S)[-1:], [l for l in s.splitlines() if 'spill' in l][-8:-7])"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --streams 4096 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RESULT $1 $2', round(d['value']), round(d['roofline']['k1_ms_per_step'],2))"
done
python -c "
from espflix_b200 import build; build.build_cuda(force=True)"
