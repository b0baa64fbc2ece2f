#!/bin/bash
# round-2 job 13: does K1a's time per stream depend on the footprint (streams per GPU)?
mkdir -p gpurun_out
for s in 512 1024 2048 4096 8192; do
  timeout 300 python bench.py --streams $s --steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts --no-verify > gpurun_out/j13_$s.json 2> gpurun_out/j13_$s.err
  python - $s <<'PY'
import json,sys
s=sys.argv[1]
try:
    d=json.load(open('gpurun_out/j13_%s.json'%s)); st=d['stages_ms']; n=int(s)
    print(s,'value %.0f'%d['value'],'k0 %.3f k1a %.3f k1b %.3f'%(st['k0_index'],st['k1a_parse'],st['k1b_recon_x12']),'per 4096 streams: k1a %.3f k1b %.3f'%(st['k1a_parse']*4096/n, st['k1b_recon_x12']*4096/n))
except Exception as e: print(s,'FAILED',e)
PY
done
