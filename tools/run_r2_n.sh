#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dbg_gemm16.py check 2>&1 | tail -8
timeout 300 python tools/dbg_gemm16.py time 2>&1 | tail -26
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "linear" 2>&1 | grep -v "UserWarning\|run_backward" | tail -5 | cut -c1-300
