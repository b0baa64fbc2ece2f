#!/bin/bash
# tools/sweep_k1.sh — rebuild libespflix_b200.so with different K1 launch shapes and time each on the GPU
# (run under gpurun; results in gpurun_out/sweep_k1.txt). Usage: tools/sweep_k1.sh "<nvcc defs>" ...
out=gpurun_out/sweep_k1.txt; : > $out
for defs in "$@"; do
  EF_NVCC_DEFS="$defs" python -c "
from espflix_b200 import build
build.build_cuda(force=True)" >/dev/null 2>&1
  EF_VERBOSE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > /tmp/sweep.json 2> /tmp/sweep.err
  python - "$defs" <<'PY' >> $out
import json, sys
try:
    d = json.load(open("/tmp/sweep.json"))
    print("%-60s value %.0f ms/step %.2f k1 %.2f e2e %.0f" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["k1_ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("%-60s FAILED %s" % (sys.argv[1], e))
PY
  grep "resident" /tmp/sweep.err | head -1 >> $out
done
python -c "
from espflix_b200 import build
build.build_cuda(force=True)" >/dev/null 2>&1
cat $out
