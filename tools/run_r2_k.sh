#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "linear" 2>&1 | grep -v "UserWarning\|run_backward" | tail -12 > gpurun_out/r2k_pytest_linear.log
cat gpurun_out/r2k_pytest_linear.log | cut -c1-700
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "wide or cfg4" 2>&1 | grep -v "UserWarning\|run_backward" | tail -6
python tools/prof_cfg4_step.py 2>&1 | tail -32
