#!/bin/bash
# One GPU call (round 1, session 2): default-path regression (tests + bench), then the opt-in f8 GEMM mode (tests, bench, one ncu capture).
# usage (from the repo root on the GPU box):  bash profiles/run_r01c.sh
O=gpurun_out/r01c
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 420 python -m pytest tests -m gpu -x -q > $O/tests_default.log 2>&1; echo "tests_default rc=$?" >> $O/status.txt
timeout 360 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench_default rc=$?" >> $O/status.txt
DSB_F8_TESTS=1 timeout 300 python -m pytest tests -m gpu -q -s -k "f8" > $O/tests_f8.log 2>&1; echo "tests_f8 rc=$?" >> $O/status.txt
timeout 360 python bench.py --precision fp16f8 --no_cpu_baseline > $O/bench_f8.json 2> $O/bench_f8.err; echo "bench_f8 rc=$?" >> $O/status.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 2 -f -o $O/f8_gemm \
    python bench.py --precision fp16f8 --steps 1 --warmup 3 --no_extras --no_cpu_baseline > $O/ncu_f8.log 2>&1; echo "ncu_f8 rc=$?" >> $O/status.txt
cat $O/status.txt
tail -3 $O/tests_default.log; tail -25 $O/tests_f8.log; cat $O/bench_default.json $O/bench_f8.json | cut -c1-1500
