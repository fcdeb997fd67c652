#!/bin/bash
# half-empty e4m3 channel blocks skipped (C = 192, 576, 64), head convolutions as BN = 32 pair / row-reuse launches: correctness, A/B
O=gpurun_out/r02aa
mkdir -p $O; rm -f $O/status.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $O/tests_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 $O/tests_kernels.log)" >> $O/status.txt
if grep -q "passed" $O/tests_kernels.log && ! grep -q "failed" $O/tests_kernels.log; then
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/tests_parity.log 2>&1; echo "parity rc=$? $(tail -1 $O/tests_parity.log)" >> $O/status.txt
ab() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 500 python bench.py --steps 6 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 "$@" > $O/ab_$name.json 2> $O/ab_$name.err
    echo "ab_$name rc=$? $(python -c "import json;d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'], d.get('forward_breakdown_ms'), (d.get('roofline') or {}).get('frac'))" 2>&1 | tail -1)" >> $O/status.txt
}
ab imagenet X=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab imagenet_half0 DSB_GEMM_F8_HALF=0 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab cifar X=1 --
ab ffhq X=1 -- --net ffhq --solver ipndm --num_steps 7 --batch 256
DSB_PRECISION=fp16f8 timeout 300 python profiles/profile_ops.py cifar10 > $O/profile_ops_cifar10_f8.txt 2>&1; echo "profile_ops rc=$?" >> $O/status.txt
fi
cat $O/status.txt | cut -c1-420; grep -E "^x1 +4096 +1 +(16|32) " $O/profile_ops_cifar10_f8.txt; grep -E "^FAILED|^ERROR|Error" $O/tests_kernels.log $O/tests_parity.log | head
