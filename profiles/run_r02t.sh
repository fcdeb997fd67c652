#!/bin/bash
# gn_apply_v3 round-robin chunk traversal: correctness, per-op bandwidth tables (CIFAR-10, ImageNet-64), benches
O=gpurun_out/r02t
mkdir -p $O; rm -f $O/status.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $O/tests_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 $O/tests_kernels.log)" >> $O/status.txt
DSB_PRECISION=fp16f8 timeout 300 python profiles/profile_ops.py cifar10 > $O/profile_ops_cifar10_f8.txt 2>&1; echo "profile_ops cifar rc=$?" >> $O/status.txt
DSB_PRECISION=fp16f8 timeout 300 python profiles/profile_ops.py imagenet64 > $O/profile_ops_imagenet64_f8.txt 2>&1; echo "profile_ops imagenet rc=$?" >> $O/status.txt
ab() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 500 python bench.py --steps 6 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 "$@" > $O/ab_$name.json 2> $O/ab_$name.err
    echo "ab_$name rc=$? $(python -c "import json;d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'], d.get('forward_breakdown_ms'), (d.get('roofline') or {}).get('frac'))" 2>&1 | tail -1)" >> $O/status.txt
}
ab cifar X=1 --
ab imagenet X=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoiser_parity or sampler_parity" > $O/tests_parity.log 2>&1; echo "parity rc=$? $(tail -1 $O/tests_parity.log)" >> $O/status.txt
cat $O/status.txt | cut -c1-420; sed -n 1,40p $O/profile_ops_cifar10_f8.txt; sed -n 1,60p $O/profile_ops_imagenet64_f8.txt
