#!/bin/bash
O=gpurun_out/r02f
mkdir -p $O; rm -f $O/status.txt
for il in 0 1; do for g in 4 3; do
DSB_ATTN_INTERLEAVE=$il DSB_ATTN_GROUPS=$g timeout 300 python profiles/attn_timeline.py 1024 64 6 > $O/attn3_timeline_g${g}_il${il}.txt 2> $O/err.txt; echo "timeline g$g il$il rc=$? $(head -1 $O/attn3_timeline_g${g}_il${il}.txt)" >> $O/status.txt
done; done
DSB_ATTN_INTERLEAVE=1 timeout 300 python profiles/attn_timeline.py 4096 16 8 > $O/attn3_timeline_L4096_il1.txt 2>> $O/err.txt; echo "L4096 il1 $(head -1 $O/attn3_timeline_L4096_il1.txt)" >> $O/status.txt
DSB_ATTN_INTERLEAVE=0 timeout 300 python profiles/attn_timeline.py 4096 16 8 > $O/attn3_timeline_L4096_il0.txt 2>> $O/err.txt; echo "L4096 il0 $(head -1 $O/attn3_timeline_L4096_il0.txt)" >> $O/status.txt
DSB_ATTN=1 timeout 300 python profiles/attn_timeline.py 4096 16 8 > $O/attn1_L4096.txt 2>> $O/err.txt; echo "L4096 v1 $(head -1 $O/attn1_L4096.txt)" >> $O/status.txt
DSB_ATTN=1 timeout 300 python profiles/attn_timeline.py 1024 64 6 > $O/attn1_L1024.txt 2>> $O/err.txt; echo "L1024 v1 $(head -1 $O/attn1_L1024.txt)" >> $O/status.txt
DSB_ATTN_INTERLEAVE=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k fused_attention 2>&1 | tail -1 >> $O/status.txt
cat $O/status.txt; tail -3 $O/err.txt
