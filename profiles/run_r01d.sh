#!/bin/bash
# GPU call 2 (round 1, session 2): the whole GPU suite with fp16f8 as the default precision (+ the gn_apply V2 A/B), then fp16f8 bench
# lines for BASELINE configs 2-4 with the fp16x3 leg measured in the same process.   usage: bash profiles/run_r01d.sh
O=gpurun_out/r01d
mkdir -p $O
export DSB_PRECISION=fp16f8
V2=1
DSB_GN_APPLY_V2=1 timeout 420 python -m pytest tests -m gpu -q > $O/tests_f8_v2.log 2>&1; rc=$?; echo "tests_f8_v2 rc=$rc" >> $O/status.txt
if [ $rc -ne 0 ]; then
    V2=0
    timeout 420 python -m pytest tests -m gpu -q > $O/tests_f8.log 2>&1; echo "tests_f8 (gn_apply v1) rc=$?" >> $O/status.txt
fi
timeout 300 python bench.py --precision fp16f8 --no_cpu_baseline > $O/bench_cifar10_f8.json 2> $O/bench_cifar10_f8.err; echo "bench cifar v1 rc=$?" >> $O/status.txt
DSB_GN_APPLY_V2=1 timeout 300 python bench.py --precision fp16f8 --no_cpu_baseline > $O/bench_cifar10_f8_gnv2.json 2> $O/bench_cifar10_f8_gnv2.err; echo "bench cifar v2 rc=$?" >> $O/status.txt
export DSB_GN_APPLY_V2=$V2
timeout 300 python bench.py --net ffhq --solver ipndm --num_steps 7 --batch 256 --precision fp16f8 --no_cpu_baseline > $O/bench_ffhq_f8.json 2> $O/bench_ffhq_f8.err; echo "bench ffhq (v2=$V2) rc=$?" >> $O/status.txt
timeout 400 python bench.py --net imagenet64 --solver dpm_pp --num_steps 11 --batch 256 --precision fp16f8 --no_cpu_baseline > $O/bench_imagenet64_f8.json 2> $O/bench_imagenet64_f8.err; echo "bench imagenet64 (v2=$V2) rc=$?" >> $O/status.txt
cat $O/status.txt
tail -4 $O/tests_f8_v2.log
grep -h -E "^FAILED|^ERROR" $O/tests_f8*.log | head -20
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['value'], d.get('fp16x3_same_run'), d.get('forward_breakdown_ms'), d.get('extras_error'))
except Exception as e:
    print('unreadable', e)
PY
done
