#!/bin/bash
# Round 2, third GPU call: attn2 with two S accumulators per group, pair-granularity GroupNorm partials (no gn_stats pass in ImageNet-64),
# gn_apply_v3 at 80 registers; GEMM N-tile microbenchmark + ncu of the slow shapes.
O=gpurun_out/r02d
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -s > $O/tests_gpu.log 2>&1; echo "tests_gpu rc=$?" >> $O/status.txt
ab() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 500 python bench.py --steps 8 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 "$@" > $O/ab_$name.json 2> $O/ab_$name.err
    echo "ab_$name rc=$? $(python -c "import json;d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'], d.get('forward_breakdown_ms'))" 2>&1 | tail -1)" >> $O/status.txt
}
ab imagenet_v2 X=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab imagenet_v1 DSB_ATTN_V1=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab sd15_v2 X=1 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
ab sd15_v1 DSB_ATTN_V1=1 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
ab cifar X=1 --
ab ffhq X=1 -- --net ffhq --solver ipndm --num_steps 7 --batch 256
timeout 900 python profiles/bench_gemm_tiles.py > $O/gemm_tiles.txt 2> $O/gemm_tiles.err; echo "gemm_tiles rc=$?" >> $O/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2_kernel -s 20 -c 1 -f -o $O/ncu_attn2_v1 \
    python bench.py --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256 --steps 1 --warmup 1 --no_extras --no_cpu_baseline > $O/ncu_attn2.log 2>&1; echo "ncu attn2 rc=$?" >> $O/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o $O/ncu_gemm_sd8_bn80 \
    python profiles/bench_gemm_tiles.py --only "sd 8^2" --bn 80 --mode x3 --reps 3 > $O/ncu_gemm_sd8.log 2>&1; echo "ncu gemm sd8 rc=$?" >> $O/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o $O/ncu_gemm_adm64_bn192 \
    python profiles/bench_gemm_tiles.py --only "adm 64^2" --bn 192 --mode f8 --reps 3 > $O/ncu_gemm_adm64.log 2>&1; echo "ncu gemm adm64 rc=$?" >> $O/status.txt
cat $O/status.txt
grep -E "passed|failed" $O/tests_gpu.log | tail -2
grep -E "^FAILED|^ERROR" $O/tests_gpu.log | head -20
cat $O/gemm_tiles.txt | cut -c1-170
