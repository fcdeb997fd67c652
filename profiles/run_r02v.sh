#!/bin/bash
# is the pair kernel's residual (660 vs 602 cycles per K block) operand traffic?  full kernel with the A loads / B loads switched off
O=gpurun_out/r02v
mkdir -p $O; rm -f $O/status.txt $O/gemm_ab_traffic.txt
timeout 300 python profiles/bench_gemm_tiles.py --only "cifar 32" --bn 256 --mode f8 --diag 0,8,16,24,6 > $O/gemm_ab_traffic.txt 2> $O/gemm.err
timeout 300 python profiles/bench_gemm_tiles.py --only "adm 64^2" --bn 192 --mode f8 --diag 0,8,16,24,6 >> $O/gemm_ab_traffic.txt 2>> $O/gemm.err
echo "rc=$?" >> $O/status.txt
cut -c1-200 $O/gemm_ab_traffic.txt
