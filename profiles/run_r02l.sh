#!/bin/bash
# tensor-pipe instruction cost microbenchmark + attention TMEM modes inside the ImageNet-64 / SD-v1.5 benches
O=gpurun_out/r02l
mkdir -p $O; rm -f $O/status.txt
timeout 300 profiles/bin/mma_microbench 2048 > $O/mma_microbench.txt 2> $O/mma_microbench.err; echo "mma_microbench rc=$?" >> $O/status.txt
ab() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 500 python bench.py --steps 6 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 "$@" > $O/ab_$name.json 2> $O/ab_$name.err
    echo "ab_$name rc=$? $(python -c "import json;d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'], d.get('forward_breakdown_ms'), (d.get('roofline') or {}).get('frac'))" 2>&1 | tail -1)" >> $O/status.txt
}
ab imagenet_tm0 DSB_ATTN_TMEM=0 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab imagenet_tm1 DSB_ATTN_TMEM=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab imagenet_tm2 DSB_ATTN_TMEM=2 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab sd15_tm0 DSB_ATTN_TMEM=0 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
ab sd15_tm2 DSB_ATTN_TMEM=2 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
cat $O/status.txt | cut -c1-420; cat $O/mma_microbench.txt
