#!/bin/bash
# round-2 job 20: asynchronous read-back with its export kernel on the read-back stream (frozen frame-store snapshot): parity + e2e
mkdir -p gpurun_out; : > gpurun_out/sweep_variants.txt
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_host_mirror_gpu.py tests/test_composite_gpu.py -m gpu -x -q > gpurun_out/j20_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j20_pytest.log
tail -3 gpurun_out/j20_pytest.log
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e-all > gpurun_out/j20_bench$i.json 2> gpurun_out/j20_bench$i.err; python - $i <<'PY'
import json,sys
d=json.load(open('gpurun_out/j20_bench%s.json'%sys.argv[1])); print('value %.0f e2e %.0f (%.3f ms) e2e_ts %.0f'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e_ts']['value']), d['stages_ms'], d.get('verify'))
PY
done
