#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/tests.log 2>&1; tail -n 5 gpurun_out/tests.log
timeout 400 python scripts/prof_step.py > gpurun_out/prof_step.log 2>&1; grep "total device" gpurun_out/prof_step.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; tail -n 1 gpurun_out/bench.log | cut -c1-400
