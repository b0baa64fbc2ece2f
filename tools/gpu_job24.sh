#!/bin/bash
# round-2 job 24 (2 GPUs of one box): the scaling legs as the driver launches them, after the K1 / read-back changes
mkdir -p gpurun_out
run() { # name nproc extra-args
  local name=$1 n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" > gpurun_out/j24_$name.json 2> gpurun_out/j24_$name.err
  echo "$name rc $?"; python - "$name" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/j24_%s.json'%sys.argv[1]))
    print(sys.argv[1],'value %.0f e2e %.0f e2e_all %s e2e_ts %s verify %s'%(d['value'],d['e2e']['value'],d.get('e2e_all',{}).get('value'),d.get('e2e_ts',{}).get('value'),d.get('verify')))
except Exception as e: print(sys.argv[1],'FAILED',e)
PY
}
run weak2 2 --steps 20 --warmup 3 --no-cpu
run strong2 2 --scaling strong --steps 5 --warmup 3 --no-cpu --no-e2e-all --no-e2e-ts
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/j24_ref2.json 2> gpurun_out/j24_ref2.err; echo "ref2 rc $?"; head -c 300 gpurun_out/j24_ref2.json
tail -3 gpurun_out/j24_weak2.err
