#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or wgrad" -p no:cacheprovider --timeout 30 > gpurun_out/t1.log 2>&1; tail -n 2 gpurun_out/t1.log
grep -q passed gpurun_out/t1.log && ! grep -q failed gpurun_out/t1.log || exit 1
timeout 60 python scripts/prof_gemm.py 2>&1 | tail -6
timeout 300 python -m pytest tests/test_gpu_matcher.py -x -q -p no:cacheprovider --timeout 60 > gpurun_out/t2.log 2>&1; tail -n 2 gpurun_out/t2.log
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1; tail -n 1 gpurun_out/bench_default.log | cut -c1-250; tail -n 1 gpurun_out/bench_default.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e', d['e2e'])"
