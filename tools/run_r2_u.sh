#!/bin/bash
# ncu --set full of the compact SASRec attention kernels inside a configs[2]-shaped step
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_live -c 4 -s 9 -o gpurun_out/r2_attn_live -f python tools/prof_sas_step.py > gpurun_out/r2_attn_live_ncu.log 2>&1
tail -3 gpurun_out/r2_attn_live_ncu.log
