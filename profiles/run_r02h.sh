#!/bin/bash
O=gpurun_out/r02h
mkdir -p $O; rm -f $O/status.txt $O/gemm_tiles.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $O/tests_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 $O/tests_kernels.log)" >> $O/status.txt
for sh in "cifar 32" "adm 64" "sd 32"; do timeout 300 python profiles/bench_gemm_tiles.py --only "$sh" >> $O/gemm_tiles.txt 2>> $O/gemm_tiles.err; done
timeout 300 python profiles/attn_timeline.py 1024 64 6 > $O/attn3_timeline_L1024.txt 2> $O/err.txt; echo "attn L1024 $(head -1 $O/attn3_timeline_L1024.txt)" >> $O/status.txt
timeout 300 python profiles/attn_timeline.py 4096 16 8 > $O/attn3_timeline_L4096.txt 2>> $O/err.txt; echo "attn L4096 $(head -1 $O/attn3_timeline_L4096.txt)" >> $O/status.txt
timeout 400 python bench.py --steps 8 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 --no_extras > $O/ab_cifar.json 2> $O/ab_cifar.err; echo "cifar $(python -c "import json;d=json.loads(open('$O/ab_cifar.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'])")" >> $O/status.txt
DSB_GEMM_2CTA=0 timeout 400 python bench.py --steps 8 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 --no_extras > $O/ab_cifar_nopair.json 2> $O/ab_cifar.err; echo "cifar nopair $(python -c "import json;d=json.loads(open('$O/ab_cifar_nopair.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'])")" >> $O/status.txt
cat $O/status.txt; grep -E "BN=256|BN=192|BN=224" $O/gemm_tiles.txt | cut -c1-160
