#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run kernels tests/test_gpu_kernels.py -k "assign or heads"
run matcher tests/test_gpu_matcher.py -k "golden or ragged"
for b in 4 16 32; do timeout 600 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline > gpurun_out/bench_b$b.log 2>&1; echo "bench$b rc=$?" | tee -a gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_b$b.log | cut -c1-230; done
