#!/bin/bash
# eight epilogue warps (two groups on alternate 32-column chunks), one producer warp again: correctness, short-K GEMMs, benches A/B
O=gpurun_out/r02x
mkdir -p $O; rm -f $O/status.txt $O/gemm_epi.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $O/tests_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 $O/tests_kernels.log)" >> $O/status.txt
if grep -q "passed" $O/tests_kernels.log && ! grep -q "failed" $O/tests_kernels.log; then
for g in 1 2; do
  DSB_GEMM_EPI_GROUPS=$g timeout 300 python profiles/bench_gemm_tiles.py --only "sd1x1 64^2 320->2560" --bn 256 2>> $O/gemm.err | sed "s/$/  [epi groups $g]/" >> $O/gemm_epi.txt
  DSB_GEMM_EPI_GROUPS=$g timeout 300 python profiles/bench_gemm_tiles.py --only "sd1x1 64^2 320->320" --bn 160 2>> $O/gemm.err | sed "s/$/  [epi groups $g]/" >> $O/gemm_epi.txt
  DSB_GEMM_EPI_GROUPS=$g timeout 300 python profiles/bench_gemm_tiles.py --only "cifar 32" --bn 256 --mode f8 2>> $O/gemm.err | sed "s/$/  [epi groups $g]/" >> $O/gemm_epi.txt
done; echo "gemm epi rc=$?" >> $O/status.txt
ab() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 500 python bench.py --steps 6 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 "$@" > $O/ab_$name.json 2> $O/ab_$name.err
    echo "ab_$name rc=$? $(python -c "import json;d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'], d.get('forward_breakdown_ms'), (d.get('roofline') or {}).get('frac'))" 2>&1 | tail -1)" >> $O/status.txt
}
ab sd15 X=1 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
ab sd15_epi1 DSB_GEMM_EPI_GROUPS=1 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
ab imagenet X=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab imagenet_epi1 DSB_GEMM_EPI_GROUPS=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab cifar X=1 --
ab cifar_epi1 DSB_GEMM_EPI_GROUPS=1 --
fi
cat $O/status.txt | cut -c1-420; cut -c1-200 $O/gemm_epi.txt; grep -E "^FAILED|^ERROR|Error" $O/tests_kernels.log | head -10
