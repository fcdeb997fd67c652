#!/bin/bash
# Round 2, GPU call: converged MMA / TMA warps with elected tcgen05 instructions (GEMM, pair GEMM, attn3).
O=gpurun_out/r02g
mkdir -p $O; rm -f $O/status.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > $O/tests_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 $O/tests_kernels.log)" >> $O/status.txt
for sh in "cifar 32" "adm 32" "adm 64" "sd 8" "sd 32"; do
  timeout 300 python profiles/bench_gemm_tiles.py --only "$sh" >> $O/gemm_tiles.txt 2>> $O/gemm_tiles.err
done; echo "gemm_tiles rc=$?" >> $O/status.txt
DSB_ATTN_INTERLEAVE=1 timeout 300 python profiles/attn_timeline.py 1024 64 6 > $O/attn3_timeline_L1024.txt 2> $O/err.txt; echo "attn L1024 $(head -1 $O/attn3_timeline_L1024.txt)" >> $O/status.txt
DSB_ATTN_INTERLEAVE=0 timeout 300 python profiles/attn_timeline.py 1024 64 6 > $O/attn3_timeline_L1024_il0.txt 2> $O/err.txt; echo "attn L1024 il0 $(head -1 $O/attn3_timeline_L1024_il0.txt)" >> $O/status.txt
timeout 300 python profiles/attn_timeline.py 4096 16 8 > $O/attn3_timeline_L4096.txt 2>> $O/err.txt; echo "attn L4096 $(head -1 $O/attn3_timeline_L4096.txt)" >> $O/status.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/tests_gpu.log 2>&1; echo "tests_gpu rc=$? $(tail -1 $O/tests_gpu.log)" >> $O/status.txt
ab() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
    env "${envs[@]}" timeout 500 python bench.py --steps 8 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 "$@" > $O/ab_$name.json 2> $O/ab_$name.err
    echo "ab_$name rc=$? $(python -c "import json;d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1]);print(round(d['value'],2), d['clocks']['sm_mhz'], d.get('forward_breakdown_ms'), (d.get('roofline') or {}).get('frac'))" 2>&1 | tail -1)" >> $O/status.txt
}
ab cifar X=1 --
ab ffhq X=1 -- --net ffhq --solver ipndm --num_steps 7 --batch 256
ab imagenet X=1 -- --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256
ab sd15 X=1 -- --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8
cat $O/status.txt | cut -c1-420
grep -E "fill_bn|BN=256" $O/gemm_tiles.txt | cut -c1-170
