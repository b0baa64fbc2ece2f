#!/bin/bash
# round-2 job 23 (final): whole GPU suite, smoke, the bench line as the driver runs it, the reference arm, and the captures behind profiles/r02_*
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/j23_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j23_pytest.log
grep -E "level-1|passed|failed|rc |Error" gpurun_out/j23_pytest.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/j23_bench.json 2> gpurun_out/j23_bench.err; echo "bench rc $?"; tail -2 gpurun_out/j23_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j23_bench_ref.json 2> gpurun_out/j23_bench_ref.err; echo "ref rc $?"
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e-all --no-verify --no-e2e-ts"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_parse_kernel -s 1 -c 1 -o gpurun_out/j23_k1a $B > gpurun_out/j23_ncu_k1a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_recon_kernel -s 13 -c 2 -o gpurun_out/j23_k1b $B > gpurun_out/j23_ncu_k1b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ef_composite_kernel -s 2 -c 1 -o gpurun_out/j23_k2n $B > gpurun_out/j23_ncu_k2n.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j23_launches.csv $B > gpurun_out/j23_launches.log 2>&1
ls gpurun_out/j23_* | wc -l
