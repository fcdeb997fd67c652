#!/bin/bash
# GPU call 3 (round 1, session 2): final state -- whole GPU suite with the defaults (B200Net fp16x3, gn_apply V2), the default bench line
# (`--precision auto` -> fp16f8 on the CIFAR-10 headline workload), smoke(), and the ncu launch list of the default bench command.
O=gpurun_out/r01e
mkdir -p $O
timeout 420 python -m pytest tests -m gpu -q > $O/tests_default.log 2>&1; echo "tests_default rc=$?" >> $O/status.txt
timeout 360 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench_default rc=$?" >> $O/status.txt
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 1 --no_extras --no_cpu_baseline > $O/ncu_list.log 2>&1; echo "ncu_list rc=$?" >> $O/status.txt
cat $O/status.txt; tail -3 $O/tests_default.log; grep -h -E "^FAILED|^ERROR" $O/tests_default.log | head; tail -2 $O/smoke.log; cut -c1-1200 $O/bench_default.json
