#!/bin/bash
# Last GPU seconds of the round: first hardware contact of the CTA-pair GEMM kernel (opt-in test, A/B against the single-CTA kernel).
O=gpurun_out/r01f
mkdir -p $O
DSB_PAIR_TESTS=1 timeout 110 python -m pytest tests/test_gpu_kernels.py -q -s -k "pair_kernel" > $O/tests_pair.log 2>&1; echo "tests_pair rc=$?" >> $O/status.txt
cat $O/status.txt; grep -E "pair conv|passed|failed|Error|error" $O/tests_pair.log | head -30
