#!/bin/bash
# Full GPU validation: all gpu-marked tests, smoke(), the default bench, the per-kernel step profile and an ncu launch list.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 400 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt; tail -n 2 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_default.log
timeout 200 python scripts/prof_step.py > gpurun_out/prof_step.log 2>&1; grep "total device" gpurun_out/prof_step.log
if [ "$1" = "ncu" ]; then
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch 8 --no-graph > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
fi
if [ "$1" = "attn" ] || [ "$2" = "attn" ]; then
  timeout 300 ncu --set full --clock-control none -k regex:attn_ -c 4 -o gpurun_out/attn_end python scripts/prof_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?" | tee -a gpurun_out/summary.txt
fi
