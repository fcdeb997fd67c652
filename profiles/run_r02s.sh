#!/bin/bash
# split producer warps (A: warp 0, B: warp 6), pair kernel for every shape with >= 74 pair tiles: correctness, ring rates, benches, per-GEMM table
O=gpurun_out/r02s
mkdir -p $O; rm -f $O/status.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $O/tests_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 $O/tests_kernels.log)" >> $O/status.txt
timeout 600 python profiles/bench_gemm_tiles.py --only "cifar 32" --bn 256 --mode f8 --diag 0,5,6 > $O/gemm_diag.txt 2> $O/gemm_diag.err; echo "diag rc=$?" >> $O/status.txt
timeout 300 python profiles/gemm_timeline.py --diag 0,5 > $O/gemm_timeline.txt 2> $O/gemm_timeline.err; echo "timeline rc=$?" >> $O/status.txt
ab() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 500 python bench.py --steps 6 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 "$@" > $O/ab_$name.json 2> $O/ab_$name.err
    echo "ab_$name rc=$? $(python -c "import json;d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'], d.get('forward_breakdown_ms'), (d.get('roofline') or {}).get('frac'))" 2>&1 | tail -1)" >> $O/status.txt
}
ab cifar X=1 --
ab cifar_oldpair DSB_GEMM_2CTA_MIN_PAIR_TILES=296 --
ab imagenet X=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab sd15 X=1 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
ab sd15_oldpair DSB_GEMM_2CTA_MIN_PAIR_TILES=296 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
ab ffhq X=1 -- --net ffhq --solver ipndm --num_steps 7 --batch 256
DSB_PRECISION=fp16f8 timeout 300 python profiles/profile_ops.py cifar10 > $O/profile_ops_cifar10_f8.txt 2>&1; echo "profile_ops rc=$?" >> $O/status.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/tests_parity.log 2>&1; echo "parity rc=$? $(tail -1 $O/tests_parity.log)" >> $O/status.txt
cat $O/status.txt | cut -c1-420; cut -c1-200 $O/gemm_diag.txt; grep -A1 "^==" $O/gemm_timeline.txt | cut -c1-330; tail -45 $O/profile_ops_cifar10_f8.txt
