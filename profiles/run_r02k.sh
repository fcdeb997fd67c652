#!/bin/bash
# attention with the query tile in TMEM as well (DSB_ATTN_TMEM=2): correctness, then speed against modes 0 / 1
O=gpurun_out/r02k
mkdir -p $O; rm -f $O/status.txt
DSB_ATTN_TMEM=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" -s > $O/tests_attn_tmem2.log 2>&1; echo "attn tmem2 rc=$? $(tail -1 $O/tests_attn_tmem2.log)" >> $O/status.txt
DSB_ATTN_TMEM=2 DSB_ATTN_INTERLEAVE=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" > $O/tests_attn_tmem2_il0.log 2>&1; echo "attn tmem2 il0 rc=$? $(tail -1 $O/tests_attn_tmem2_il0.log)" >> $O/status.txt
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" > $O/tests_attn_default.log 2>&1; echo "attn default(tmem1) rc=$? $(tail -1 $O/tests_attn_default.log)" >> $O/status.txt
for cfg in "1 4" "1 3" "2 3"; do set -- $cfg
  DSB_ATTN_TMEM=$1 DSB_ATTN_GROUPS=$2 timeout 200 python profiles/attn_timeline.py 1024 64 6 > $O/tl_L1024_tm$1_g$2.txt 2>> $O/err.txt; echo "tmem=$1 g=$2 L1024 $(head -1 $O/tl_L1024_tm$1_g$2.txt)" >> $O/status.txt
  DSB_ATTN_TMEM=$1 DSB_ATTN_GROUPS=$2 timeout 200 python profiles/attn_timeline.py 4096 16 8 > $O/tl_L4096_tm$1_g$2.txt 2>> $O/err.txt; echo "tmem=$1 g=$2 L4096 $(head -1 $O/tl_L4096_tm$1_g$2.txt)" >> $O/status.txt
done
DSB_ATTN_TMEM=2 DSB_ATTN_INTERLEAVE=0 timeout 200 python profiles/attn_timeline.py 1024 64 6 > $O/tl_L1024_tm2_il0.txt 2>> $O/err.txt; echo "tmem=2 il0 L1024 $(head -1 $O/tl_L1024_tm2_il0.txt)" >> $O/status.txt
cat $O/status.txt; grep -E "FAILED|Error|error" $O/tests_attn_tmem2.log | head
