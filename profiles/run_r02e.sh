#!/bin/bash
# Round 2, fourth GPU call: attn3 (four / three softmax groups, output accumulators in TMEM with lazy rescaling), parallel gn_finalize,
# measured tile cost model (640 -> 224 x 3).
O=gpurun_out/r02e
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "fused_attention" > $O/tests_attn3_g4.log 2>&1; echo "attn3 g4 rc=$? $(tail -1 $O/tests_attn3_g4.log)" >> $O/status.txt
DSB_ATTN_GROUPS=3 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "fused_attention" > $O/tests_attn3_g3.log 2>&1; echo "attn3 g3 rc=$? $(tail -1 $O/tests_attn3_g3.log)" >> $O/status.txt
timeout 1800 python -m pytest tests -m gpu -q -s > $O/tests_gpu.log 2>&1; echo "tests_gpu rc=$? $(tail -1 $O/tests_gpu.log)" >> $O/status.txt
ab() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 500 python bench.py --steps 8 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 "$@" > $O/ab_$name.json 2> $O/ab_$name.err
    echo "ab_$name rc=$? $(python -c "import json;d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'], d.get('forward_breakdown_ms'))" 2>&1 | tail -1)" >> $O/status.txt
}
ab imagenet_a3g4 X=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab imagenet_a3g3 DSB_ATTN_GROUPS=3 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab imagenet_a2 DSB_ATTN=2 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab sd15_a3g4 X=1 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
ab sd15_a3g3 DSB_ATTN_GROUPS=3 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
ab sd15_a1 DSB_ATTN=1 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
ab cifar X=1 --
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn3_kernel -s 20 -c 1 -f -o $O/ncu_attn3 \
    python bench.py --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256 --steps 1 --warmup 1 --no_extras --no_cpu_baseline > $O/ncu_attn3.log 2>&1; echo "ncu attn3 rc=$?" >> $O/status.txt
cat $O/status.txt
grep -E "^FAILED|^ERROR" $O/tests_gpu.log | head -20
grep -E "fused attention" $O/tests_attn3_g4.log | head -12
