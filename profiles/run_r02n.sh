#!/bin/bash
# packed f8-image / hi-lo stores (gn_apply instruction diet); short-K 1x1 GEMMs of the SD transformer: tile table + one ncu capture
O=gpurun_out/r02n
mkdir -p $O; rm -f $O/status.txt $O/gemm_tiles_1x1.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $O/tests_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 $O/tests_kernels.log)" >> $O/status.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "f8 or denoiser_parity or ldm_cfg" > $O/tests_parity_f8.log 2>&1; echo "parity f8 rc=$? $(tail -1 $O/tests_parity_f8.log)" >> $O/status.txt
timeout 400 python bench.py --steps 6 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 > $O/ab_cifar.json 2> $O/ab_cifar.err
echo "ab_cifar rc=$? $(python -c "import json;d=json.loads(open('$O/ab_cifar.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'], d.get('forward_breakdown_ms'), (d.get('roofline') or {}).get('frac'))" 2>&1 | tail -1)" >> $O/status.txt
timeout 600 python profiles/bench_gemm_tiles.py --only sd1x1 > $O/gemm_tiles_1x1.txt 2> $O/gemm_tiles.err; echo "gemm tiles rc=$?" >> $O/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -f -o $O/ncu_gemm_sd1x1_320_2560_f8 \
    python profiles/bench_gemm_tiles.py --only "sd1x1 64^2 320->2560" --bn 256 --mode f8 --reps 2 > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" >> $O/status.txt
cat $O/status.txt | cut -c1-400; cat $O/gemm_tiles_1x1.txt | cut -c1-170
