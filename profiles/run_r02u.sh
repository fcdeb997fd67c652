#!/bin/bash
# two K blocks per ring stage (DSB_GEMM_GROUP=2 default): correctness, BN=192 / BN=128 shapes, benches with A/B against group 1
O=gpurun_out/r02u
mkdir -p $O; rm -f $O/status.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $O/tests_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 $O/tests_kernels.log)" >> $O/status.txt
for g in 1 2; do
  DSB_GEMM_GROUP=$g timeout 300 python profiles/bench_gemm_tiles.py --only "adm 64^2" --bn 192 --diag 0 2>> $O/gemm.err | sed "s/$/  [group $g]/" >> $O/gemm_group.txt
  DSB_GEMM_GROUP=$g timeout 300 python profiles/bench_gemm_tiles.py --only "cifar 32" --bn 256 --mode f8 --diag 0 2>> $O/gemm.err | sed "s/$/  [group $g]/" >> $O/gemm_group.txt
  DSB_GEMM_GROUP=$g timeout 300 python profiles/bench_gemm_tiles.py --only "sd 32^2" --bn 224 --mode f8 --diag 0 2>> $O/gemm.err | sed "s/$/  [group $g]/" >> $O/gemm_group.txt
done; echo "gemm group rc=$?" >> $O/status.txt
ab() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 500 python bench.py --steps 6 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 "$@" > $O/ab_$name.json 2> $O/ab_$name.err
    echo "ab_$name rc=$? $(python -c "import json;d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'], d.get('forward_breakdown_ms'), (d.get('roofline') or {}).get('frac'))" 2>&1 | tail -1)" >> $O/status.txt
}
ab imagenet X=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab imagenet_g1 DSB_GEMM_GROUP=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab cifar X=1 --
ab cifar_g1 DSB_GEMM_GROUP=1 --
ab ffhq X=1 -- --net ffhq --solver ipndm --num_steps 7 --batch 256
ab sd15 X=1 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoiser_parity or sampler_parity or config" > $O/tests_parity.log 2>&1; echo "parity rc=$? $(tail -1 $O/tests_parity.log)" >> $O/status.txt
cat $O/status.txt | cut -c1-420; cut -c1-200 $O/gemm_group.txt
