#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run kernels tests/test_gpu_kernels.py
run matcher tests/test_gpu_matcher.py
timeout 600 python bench.py --steps 10 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/bench_b4.log 2>&1; echo "bench4 rc=$?" | tee -a gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_b4.log | cut -c1-260
timeout 600 python bench.py --steps 10 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/bench_b16.log 2>&1; echo "bench16 rc=$?" | tee -a gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_b16.log | cut -c1-260
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch 4 --no-cpu-baseline --no-kernel-timing --no-graph > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
