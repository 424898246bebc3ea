#!/bin/bash
# Full GPU validation: all gpu-marked tests, smoke(), the default bench and the reference (CPU) arm.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt; tail -n 4 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_reference.log 2>&1; echo "bench_ref rc=$?" | tee -a gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_reference.log
