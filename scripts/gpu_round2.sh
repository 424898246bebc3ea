#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run attn tests/test_gpu_kernels.py -k attention
run matcher tests/test_gpu_matcher.py
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --batch 4 > gpurun_out/bench_b4.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt; tail -2 gpurun_out/bench_b4.log
timeout 600 python bench.py --steps 5 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/bench_b8.log 2>&1; tail -1 gpurun_out/bench_b8.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch 4 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
