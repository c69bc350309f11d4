#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm16_kernel -c 1 -s 1 -o gpurun_out/r2_gemm16 -f python tools/dbg_gemm16.py prof > gpurun_out/r2_gemm16_ncu.log 2>&1
tail -3 gpurun_out/r2_gemm16_ncu.log
