#!/bin/bash
# P-in-TMEM attention variant: correctness on every attention shape, then speed against the shared-memory P variant
O=gpurun_out/r02j
mkdir -p $O; rm -f $O/status.txt
DSB_ATTN_PTMEM=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" -s > $O/tests_attn_ptmem.log 2>&1; echo "attn ptmem rc=$? $(tail -1 $O/tests_attn_ptmem.log)" >> $O/status.txt
DSB_ATTN_PTMEM=1 DSB_ATTN_INTERLEAVE=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" -s > $O/tests_attn_ptmem_il0.log 2>&1; echo "attn ptmem il0 rc=$? $(tail -1 $O/tests_attn_ptmem_il0.log)" >> $O/status.txt
for pt in 0 1; do for il in 1 0; do
  DSB_ATTN_PTMEM=$pt DSB_ATTN_INTERLEAVE=$il timeout 200 python profiles/attn_timeline.py 1024 64 6 > $O/tl_L1024_pt${pt}_il${il}.txt 2>> $O/err.txt; echo "pt=$pt il=$il L1024 $(head -1 $O/tl_L1024_pt${pt}_il${il}.txt)" >> $O/status.txt
  DSB_ATTN_PTMEM=$pt DSB_ATTN_INTERLEAVE=$il timeout 200 python profiles/attn_timeline.py 4096 16 8 > $O/tl_L4096_pt${pt}_il${il}.txt 2>> $O/err.txt; echo "pt=$pt il=$il L4096 $(head -1 $O/tl_L4096_pt${pt}_il${il}.txt)" >> $O/status.txt
done; done
DSB_ATTN_PTMEM=1 DSB_ATTN_GROUPS=3 timeout 200 python profiles/attn_timeline.py 1024 64 6 > $O/tl_L1024_pt1_g3.txt 2>> $O/err.txt; echo "pt=1 g3 L1024 $(head -1 $O/tl_L1024_pt1_g3.txt)" >> $O/status.txt
cat $O/status.txt; grep -E "FAILED|Error|error" $O/tests_attn_ptmem.log | head
