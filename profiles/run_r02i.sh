#!/bin/bash
O=gpurun_out/r02i
mkdir -p $O; rm -f $O/status.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" -s > $O/tests_attn.log 2>&1; echo "attn rc=$? $(tail -1 $O/tests_attn.log)" >> $O/status.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "clip" -s > $O/tests_clip.log 2>&1; echo "clip rc=$? $(tail -1 $O/tests_clip.log)" >> $O/status.txt
cat $O/status.txt; grep -E "causal=True|last_hidden|Error|error" $O/tests_attn.log $O/tests_clip.log | head -30
