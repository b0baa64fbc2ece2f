#!/bin/bash
# round-2 job 28: compute-sanitizer (memcheck, then racecheck on shared memory) over the smoke decode and one coverage stream
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j28_memcheck.log 2>&1; echo "memcheck rc $?"; grep -E "ERROR SUMMARY|smoke ok|Invalid|out of bounds" gpurun_out/j28_memcheck.log | head -5
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j28_racecheck.log 2>&1; echo "racecheck rc $?"; grep -E "RACECHECK SUMMARY|smoke ok|hazard" gpurun_out/j28_racecheck.log | head -8
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_decode_gpu.py -m gpu -q -x -k "reference_fixture_ts_path or truncated" > gpurun_out/j28_memcheck2.log 2>&1; echo "memcheck2 rc $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/j28_memcheck2.log | head -5
