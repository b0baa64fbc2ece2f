#!/bin/bash
# round-2 job 6 (8 GPUs of one box): weak and strong scaling with NUMA-bound ranks, and the same without the binding
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/j6_topo.txt 2>&1
run() { # name nproc extra-args
  local name=$1 n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" > gpurun_out/j6_$name.json 2> gpurun_out/j6_$name.err
  echo "$name rc $?"; python - "$name" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/j6_%s.json'%sys.argv[1]))
    print(sys.argv[1],'value %.0f e2e %.0f e2e_all %s e2e_ts %s link %s'%(d['value'],d['e2e']['value'],d.get('e2e_all',{}).get('value'),d.get('e2e_ts',{}).get('value'),json.dumps(d['host_link']['all_ranks_sum_gbs'])))
except Exception as e: print(sys.argv[1],'FAILED',e)
PY
}
run weak8 8 --steps 20 --warmup 3 --no-cpu
run weak8_nonuma 8 --steps 20 --warmup 3 --no-cpu --no-numa --no-e2e-all --no-e2e-ts
run weak4 4 --steps 20 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts
run weak2 2 --steps 20 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts
run strong8 8 --scaling strong --steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts
run strong4 4 --scaling strong --steps 10 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts
run strong2 2 --scaling strong --steps 5 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts
tail -3 gpurun_out/j6_weak8.err
