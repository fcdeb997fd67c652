#!/bin/bash
# Round 2, first GPU call: parity suite incl. the new BASELINE-config / drop-in tests, hardware contact for the opt-in tests, memcheck of the
# smoke path, the default bench (configs 2-5 + GPU-eager comparator), the A/Bs that decide defaults, and an ncu launch list.
# usage (repo root on the GPU box):  bash profiles/run_r02b.sh
O=gpurun_out/r02b
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -x -s > $O/tests_gpu.log 2>&1; echo "tests_gpu rc=$?" >> $O/status.txt
DSB_PAIR_TESTS=1 DSB_LDM_F8_LINEAR_TESTS=1 DSB_VAE_TESTS=1 timeout 600 python -m pytest tests -m gpu -q -s \
    -k "pair_kernel or f8_image or f8_linear or vae_decoder" > $O/tests_optin.log 2>&1; echo "tests_optin rc=$?" >> $O/status.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?" >> $O/status.txt
timeout 1200 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench_default rc=$?" >> $O/status.txt
DSB_GEMM_2CTA=1 timeout 300 python bench.py --steps 10 --warmup 3 --no_extras --no_cpu_baseline > $O/bench_pair.json 2> $O/bench_pair.err; echo "bench_pair rc=$?" >> $O/status.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no_extras --no_cpu_baseline > $O/bench_nopair.json 2> $O/bench_nopair.err; echo "bench_nopair rc=$?" >> $O/status.txt
# FFHQ: f8 only in the blocks with >= 256 channels
timeout 400 python bench.py --net ffhq --solver ipndm --num_steps 7 --batch 256 --precision fp16f8 --f8_min_channels 256 --no_cpu_baseline --all_configs 0 --gpu_eager 0 \
    > $O/bench_ffhq_f8_min256.json 2> $O/bench_ffhq_f8_min256.err; echo "bench_ffhq f8>=256 rc=$?" >> $O/status.txt
for mode in "fp16f8 0" "fp16f8 1"; do
    set -- $mode
    DSB_LDM_F8_LINEAR=$2 timeout 600 python bench.py --net sd15 --solver amed_dpm_pp --num_steps 4 --batch 8 --precision $1 --no_cpu_baseline --all_configs 0 --gpu_eager 0 \
        > $O/bench_sd15_$1_lin$2.json 2> $O/bench_sd15_$1_lin$2.err; echo "bench_sd15 $1 f8_linear=$2 rc=$?" >> $O/status.txt
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no_extras --no_cpu_baseline > $O/ncu_bench.log 2>&1; echo "ncu launch list rc=$?" >> $O/status.txt
cat $O/status.txt
tail -5 $O/tests_gpu.log
grep -E "passed|failed" $O/tests_optin.log | tail -3
tail -3 $O/sanitizer_memcheck_smoke.log
for f in $O/bench_*.json; do echo "$f: $(python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(round(d['value'],1), d.get('precision'), 'e2e', d.get('e2e',{}).get('value'), 'rl', (d.get('roofline') or {}).get('frac'), 'fwd', (d.get('roofline') or {}).get('forward_ms_back_to_back'), d.get('fp16x3_same_run'))
for c in d.get('configs', []): print('   cfg', c.get('id'), c.get('value'), c.get('precision'), (c.get('roofline') or {}).get('frac'), c.get('error'), {k: (v.get('value') if isinstance(v, dict) else v) for k, v in (c.get('gpu_eager') or {}).items() if k in ('default','fp16','fp32','error')})
print('   eager', {k: (v.get('value') if isinstance(v, dict) else v) for k, v in (d.get('gpu_eager') or {}).items() if k in ('default','fp16','fp32','error','ratio_vs_default','ratio_vs_fp16')})
" 2>&1 | cut -c1-600)"; done
