#!/bin/bash
# row reuse in the pair kernel (DSB_GEMM_RR=1 default): correctness, A/B on the conv shapes and on the benches
O=gpurun_out/r02w
mkdir -p $O; rm -f $O/status.txt $O/gemm_rr.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "pair or conv" > $O/tests_conv.log 2>&1; echo "conv tests rc=$? $(tail -1 $O/tests_conv.log)" >> $O/status.txt
if grep -q "passed" $O/tests_conv.log && ! grep -q "failed" $O/tests_conv.log; then
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $O/tests_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 $O/tests_kernels.log)" >> $O/status.txt
for rr in 0 1; do
  DSB_GEMM_RR=$rr timeout 300 python profiles/bench_gemm_tiles.py --only "cifar 32" --bn 256 --diag 0 2>> $O/gemm.err | grep "pair=1" | sed "s/$/  [rr $rr]/" >> $O/gemm_rr.txt
  DSB_GEMM_RR=$rr timeout 300 python profiles/bench_gemm_tiles.py --only "adm 64^2" --bn 192 --diag 0 2>> $O/gemm.err | grep "pair=1" | sed "s/$/  [rr $rr]/" >> $O/gemm_rr.txt
  DSB_GEMM_RR=$rr timeout 300 python profiles/bench_gemm_tiles.py --only "adm 16^2" --mode f8 --diag 0 2>> $O/gemm.err | grep "pair=1" | sed "s/$/  [rr $rr]/" >> $O/gemm_rr.txt
done; echo "gemm rr rc=$?" >> $O/status.txt
ab() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 500 python bench.py --steps 6 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 "$@" > $O/ab_$name.json 2> $O/ab_$name.err
    echo "ab_$name rc=$? $(python -c "import json;d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'], d.get('forward_breakdown_ms'), (d.get('roofline') or {}).get('frac'))" 2>&1 | tail -1)" >> $O/status.txt
}
ab cifar X=1 --
ab cifar_rr0 DSB_GEMM_RR=0 --
ab imagenet X=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab imagenet_rr0 DSB_GEMM_RR=0 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoiser_parity or sampler_parity or config" > $O/tests_parity.log 2>&1; echo "parity rc=$? $(tail -1 $O/tests_parity.log)" >> $O/status.txt
fi
cat $O/status.txt | cut -c1-420; cut -c1-200 $O/gemm_rr.txt; grep -E "^FAILED|^ERROR|Error|error" $O/tests_conv.log | head -10
