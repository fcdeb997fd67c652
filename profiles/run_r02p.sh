#!/bin/bash
# what bounds the conv GEMM: DSB_GEMM_DIAG modes (no MMA / no TMA / no epilogue) on the CIFAR-10 256->256 32x32 shape, single and pair
# kernels, f8 and x3, with the SM clock under each load; ncu --set full of both kernels on that shape
O=gpurun_out/r02p
mkdir -p $O; rm -f $O/status.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $O/tests_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 $O/tests_kernels.log)" >> $O/status.txt
timeout 900 python profiles/bench_gemm_tiles.py --only "cifar 32" --bn 256 --diag 0,1,2,4,5,6 > $O/gemm_diag_cifar.txt 2> $O/gemm_diag.err; echo "diag rc=$?" >> $O/status.txt
timeout 600 python profiles/bench_gemm_tiles.py --only "adm 64^2" --bn 192 --diag 0,1,2,4,5,6 >> $O/gemm_diag_cifar.txt 2>> $O/gemm_diag.err; echo "diag adm rc=$?" >> $O/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair_kernel -s 1 -c 1 -f -o $O/ncu_gemm_pair_cifar_f8 \
    python profiles/bench_gemm_tiles.py --only "cifar 32" --bn 256 --mode f8 --reps 2 > $O/ncu_pair.log 2>&1; echo "ncu pair rc=$?" >> $O/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -f -o $O/ncu_gemm_single_cifar_f8 \
    python profiles/bench_gemm_tiles.py --only "cifar 32" --bn 256 --mode f8 --reps 2 > $O/ncu_single.log 2>&1; echo "ncu single rc=$?" >> $O/status.txt
cat $O/status.txt | cut -c1-300; cat $O/gemm_diag_cifar.txt | cut -c1-200
