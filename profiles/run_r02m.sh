#!/bin/bash
# ncu full captures of gn_apply_v3 (CIFAR-10) and attn3 mode 2 (ImageNet-64); per-op / per-shape profile of the SD-v1.5 forward
O=gpurun_out/r02m
mkdir -p $O; rm -f $O/status.txt
timeout 600 python profiles/profile_sd15.py fp16f8 > $O/profile_sd15_fp16f8.txt 2> $O/profile_sd15.err; echo "profile_sd15 rc=$?" >> $O/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_apply_v3 -s 40 -c 2 -f -o $O/ncu_gn_apply_v3_cifar \
    python bench.py --steps 1 --warmup 1 --no_extras --no_cpu_baseline --all_configs 0 --gpu_eager 0 > $O/ncu_gn.log 2>&1; echo "ncu gn rc=$?" >> $O/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn3_kernel -s 6 -c 2 -f -o $O/ncu_attn3_tmem2_imagenet \
    python bench.py --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256 --steps 1 --warmup 1 --no_extras --no_cpu_baseline --all_configs 0 --gpu_eager 0 > $O/ncu_attn.log 2>&1; echo "ncu attn rc=$?" >> $O/status.txt
cat $O/status.txt; head -20 $O/profile_sd15_fp16f8.txt | cut -c1-250; grep -A46 "GEMM shapes" $O/profile_sd15_fp16f8.txt | cut -c1-200
