#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | grep -v "UserWarning\|run_backward" | tail -30 > gpurun_out/r2l_pytest_attn.log
cat gpurun_out/r2l_pytest_attn.log | cut -c1-600
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "wide or cfg4" 2>&1 | grep -v "UserWarning\|run_backward" | tail -6 | cut -c1-400
python tools/prof_cfg4_step.py 2>&1 | head -12
