#!/bin/bash
# round-2 job 25 (8 GPUs of one box): weak scaling with the final kernels (NUMA-bound ranks)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/j25_weak8.json 2> gpurun_out/j25_weak8.err
echo "weak8 rc $?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/j25_weak8.json'))
    print('weak8 value %.0f e2e %.0f e2e_all %s e2e_ts %s link %s verify %s'%(d['value'],d['e2e']['value'],d.get('e2e_all',{}).get('value'),d.get('e2e_ts',{}).get('value'),json.dumps(d['host_link']['all_ranks_sum_gbs']),d.get('verify')))
except Exception as e: print('FAILED',e)
PY
tail -2 gpurun_out/j25_weak8.err
