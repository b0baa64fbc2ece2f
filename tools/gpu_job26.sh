#!/bin/bash
# round-2 job 26: the GPU suite and smoke on the final HEAD (after the read-back ordering guards)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/j26_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j26_pytest.log
tail -3 gpurun_out/j26_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --no-cpu > gpurun_out/j26_bench.json 2> gpurun_out/j26_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/j26_bench.json')); print('value %.0f e2e %.0f steps %d clocks %s'%(d['value'],d['e2e']['value'],d['steps'],d['clocks']))
PY
