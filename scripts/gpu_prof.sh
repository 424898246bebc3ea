#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 6 -c 2 -o gpurun_out/attn_v3 python scripts/prof_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu rc=$?"
