#!/bin/bash
# final validation of the round: whole GPU suite, the driver's bench command, launch list + ncu --set full of the dominant kernel, memcheck
O=gpurun_out/r02y
mkdir -p $O; rm -f $O/status.txt
timeout 1500 python -m pytest tests -m gpu -q -x > $O/tests_gpu.log 2>&1; echo "tests_gpu rc=$? $(tail -1 $O/tests_gpu.log)" >> $O/status.txt
timeout 1200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench_default rc=$? $(python -c "import json;d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1]);print(round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],4), [(c.get('workload','')[:28], round(c.get('value') or 0,2)) for c in d.get('configs',[])])" 2>&1 | tail -1)" >> $O/status.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --all_configs 0 --gpu_eager 0 --no_extras > $O/launches_bench.log 2>&1; echo "ncu launch list rc=$?" >> $O/status.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair_kernel -s 1 -c 1 -f -o $O/ncu_gemm_pair_rr_cifar_f8 \
    python profiles/bench_gemm_tiles.py --only "cifar 32" --bn 256 --mode f8 --reps 2 > $O/ncu_pair.log 2>&1; echo "ncu pair rc=$?" >> $O/status.txt
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "test_conv_pair_kernel and 3-16-16" > $O/sanitizer_pair_rr.log 2>&1; echo "sanitizer rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $O/sanitizer_pair_rr.log | tr '\n' ' ')" >> $O/status.txt
cat $O/status.txt | cut -c1-600; tail -3 $O/bench_default.err | cut -c1-300
