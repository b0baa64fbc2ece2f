#!/bin/bash
# round-2 job 4: whole GPU suite (K2 v6, batched level-1 mirror, audio), bench line, strong-scaling fit at N=1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/j4_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j4_pytest.log
grep -E "level-1|passed|failed|rc " gpurun_out/j4_pytest.log | tail -8
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/j4_bench.json 2> gpurun_out/j4_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/j4_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'e2e_all',d.get('e2e_all',{}).get('value'),'k2',d['composite']['ntsc']['frac'],d['composite']['pal']['frac'],d['stages_ms'])
PY
timeout 1200 python bench.py --scaling strong --steps 5 --warmup 3 --no-cpu --no-e2e-all > gpurun_out/j4_strong1.json 2> gpurun_out/j4_strong1.err; echo "strong rc $?"; tail -3 gpurun_out/j4_strong1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/j4_strong1.json')); print('strong N=1 value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],d['config']['workload'])
except Exception as e: print('strong failed',e)
PY
