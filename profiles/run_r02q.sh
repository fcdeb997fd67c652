#!/bin/bash
# what bounds the operand feed of the conv GEMM (r02p: ~930 cycles per ring stage whatever the B bytes): A-only / B-only / unshifted taps /
# L2-hot A tile / ring depth variants of the feed-only mode, and the ring timeline of CTA 0
O=gpurun_out/r02q
mkdir -p $O; rm -f $O/status.txt
timeout 600 python profiles/bench_gemm_tiles.py --only "cifar 32" --bn 256 --mode f8 --diag 0,5,13,21,37,69,101 > $O/gemm_diag2_cifar.txt 2> $O/gemm_diag2.err; echo "diag2 rc=$?" >> $O/status.txt
for st in 2 3; do DSB_GEMM_STAGES=$st timeout 300 python profiles/bench_gemm_tiles.py --only "cifar 32" --bn 256 --mode f8 --diag 0,5 2>> $O/gemm_diag2.err | sed "s/$/  [stages<=$st]/" >> $O/gemm_diag2_cifar.txt; done; echo "stages rc=$?" >> $O/status.txt
timeout 300 python profiles/gemm_timeline.py --diag 0,5 > $O/gemm_timeline.txt 2> $O/gemm_timeline.err; echo "timeline rc=$?" >> $O/status.txt
cat $O/status.txt; cut -c1-210 $O/gemm_diag2_cifar.txt; grep -A1 "^==" $O/gemm_timeline.txt | cut -c1-400; tail -3 $O/gemm_timeline.err
