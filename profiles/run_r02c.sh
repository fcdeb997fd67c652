#!/bin/bash
# Round 2, second GPU call: full parity suite with the round-2 kernels on by default (attn2, gn_apply_v3 + coefficient tables, CUDA-graph
# replay, CTA-pair GEMM, AMED predictor kernel, f8_linear), default bench, and A/Bs of each switch on the same box.
O=gpurun_out/r02c
mkdir -p $O
timeout 2700 python -m pytest tests -m gpu -q -s > $O/tests_gpu.log 2>&1; echo "tests_gpu rc=$?" >> $O/status.txt
timeout 1500 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench_default rc=$?" >> $O/status.txt
ab() { # name, env..., -- bench args
    name=$1; shift
    envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 400 python bench.py --steps 8 --warmup 3 --no_extras --no_cpu_baseline "$@" > $O/ab_$name.json 2> $O/ab_$name.err
    echo "ab_$name rc=$? $(python -c "import json;d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'])" 2>&1 | tail -1)" >> $O/status.txt
}
ab cifar_all_on X=1 --
ab cifar_graph_off DSB_CUDA_GRAPH=0 --
ab cifar_gncoef_off DSB_GN_COEF=0 --
ab cifar_pair_off DSB_GEMM_2CTA=0 --
ab cifar_all_on_again X=1 --
ab imagenet_all_on X=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab imagenet_attn_v1 DSB_ATTN_V1=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab sd15_x3_graph_on X=1 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8 --precision fp16x3
ab sd15_x3_graph_off DSB_CUDA_GRAPH=0 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8 --precision fp16x3
ab sd15_x3_attn_v1 DSB_ATTN_V1=1 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8 --precision fp16x3
# SD-v1.5 in fp16f8 (+ f8_linear) with the fp16x3 run of the same latents in the same process (final-image difference)
timeout 900 python bench.py --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8 --precision fp16f8 --steps 6 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 \
    > $O/bench_sd15_fp16f8.json 2> $O/bench_sd15_fp16f8.err; echo "bench_sd15_f8 rc=$?" >> $O/status.txt
# ncu: full-section captures of the new kernels inside the ImageNet-64 bench command (2 launches each)
for k in attn2_kernel gn_apply_v3_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 2 -f -o $O/ncu_$k \
      python bench.py --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256 --steps 1 --warmup 1 --no_extras --no_cpu_baseline > $O/ncu_$k.log 2>&1
  echo "ncu $k rc=$?" >> $O/status.txt
done
cat $O/status.txt
grep -E "passed|failed" $O/tests_gpu.log | tail -3
grep -E "^FAILED|^ERROR" $O/tests_gpu.log | head -30
python - <<'PY'
import json
for f in ('bench_default','bench_sd15_fp16f8'):
    try:
        d=json.loads(open('gpurun_out/r02c/%s.json'%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    rl=d.get('roofline') or {}
    print(f, round(d['value'],1), d.get('precision'), 'e2e', round((d.get('e2e') or {}).get('value',0),1), 'frac', rl.get('frac'), 'fwd_b2b', rl.get('forward_ms_back_to_back'), 'ops', rl.get('all_ops_ms_per_forward'), d.get('fp16x3_same_run'), d.get('extras_error'))
    print('   breakdown', d.get('forward_breakdown_ms'))
    for c in d.get('configs', []):
        r=c.get('roofline') or {}
        print('   cfg', c.get('id'), c.get('value'), c.get('precision'), c.get('f8_min_channels'), 'frac', r.get('frac'), 'fwd', r.get('forward_ms_back_to_back'), c.get('error'), {k:(round(v['value'],1) if isinstance(v,dict) else v) for k,v in (c.get('gpu_eager') or {}).items() if k in ('default','fp16','fp32','error')}, c.get('gits'))
        print('       ', c.get('forward_breakdown_ms'))
    print('   eager', {k:(round(v['value'],1) if isinstance(v,dict) else v) for k,v in (d.get('gpu_eager') or {}).items() if k in ('default','fp16','fp32','error')})
PY
