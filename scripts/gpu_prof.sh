#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -k attention -x -q -p no:cacheprovider --timeout 30 > gpurun_out/attn.log 2>&1; tail -n 2 gpurun_out/attn.log
grep -q passed gpurun_out/attn.log && ! grep -q failed gpurun_out/attn.log || exit 1
timeout 60 python scripts/prof_attn.py > gpurun_out/prof_attn.log 2>&1; tail -n 8 gpurun_out/prof_attn.log
PB=32 timeout 60 python scripts/prof_attn.py > gpurun_out/prof_attn32.log 2>&1; tail -n 1 gpurun_out/prof_attn32.log
