#!/bin/bash
# Round-2 end-of-round GPU validation (one gpurun call): all gpu tests, smoke(), the default bench and the N=1024 config,
# the per-kernel step table, the ncu launch list of the bench command, one `--set full` capture of the attention backward
# (DRAM bytes per launch for bench.py's roofline.traffic), Sinkhorn / GlueStick timings.  Outputs under gpurun_out/.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
rm -f gpurun_out/parity_report.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt; tail -c 400 gpurun_out/bench_default.json
timeout 400 python bench.py --kpts 1024 --no-cpu-baseline > gpurun_out/bench_n1024.json 2> gpurun_out/bench_n1024.err; echo "bench n1024 rc=$?" | tee -a gpurun_out/summary.txt; tail -c 300 gpurun_out/bench_n1024.json
timeout 300 python scripts/prof_step.py > gpurun_out/prof_step.log 2>&1; grep "total device" gpurun_out/prof_step.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 1000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:attn_bwd|dq_convert" -s 12 -c 9 -o gpurun_out/r02_attn_bwd -f python scripts/prof_step.py > gpurun_out/ncu_attn_bwd.log 2>&1; echo "ncu attn_bwd rc=$?" | tee -a gpurun_out/summary.txt
timeout 200 python scripts/prof_sinkhorn.py > gpurun_out/prof_sinkhorn.log 2>&1; cat gpurun_out/prof_sinkhorn.log
PB=8 timeout 400 python scripts/prof_gluestick.py > gpurun_out/prof_gluestick.log 2>&1; tail -8 gpurun_out/prof_gluestick.log
ls -la gpurun_out/*.ncu-rep
