#!/bin/bash
# ncu --set full captures of the round-2 hot kernels inside one eager training step at the bench batch size
# (32 pairs = 64 sequences): the persistent GEMM variants, the fused assignment kernels, the attention kernels.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_bf16_kernel -s 40 -c 10 -o gpurun_out/r02_gemm -f python scripts/prof_step.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 300 $NCU -k regex:assign_tc_kernel -c 3 -o gpurun_out/r02_assign -f python scripts/prof_step.py > gpurun_out/ncu_assign.log 2>&1; echo "ncu assign rc=$?"
timeout 300 $NCU -k regex:attn_ -c 4 -o gpurun_out/r02_attn -f python scripts/prof_step.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
# per-launch DRAM traffic of the attention backward at the bench batch size (for bench.py's roofline.traffic)
timeout 300 $NCU -k regex:attn_bwd -c 3 -o gpurun_out/r02_attn_bwd -f python scripts/prof_step.py > gpurun_out/ncu_attn_bwd.log 2>&1; echo "ncu attn_bwd rc=$?"
