#!/bin/bash
# Runs on the GPU box (via gpurun): per-group pytest runs, each in its own process with a hard timeout,
# so that a hung kernel in one group cannot take the others down. Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run gemm tests/test_gpu_kernels.py -k gemm
run attn tests/test_gpu_kernels.py -k attention
run misc tests/test_gpu_kernels.py -k "not gemm and not attention"
run matcher tests/test_gpu_matcher.py
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/smoke.log
