#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "linear" 2>&1 | grep -v "UserWarning\|run_backward" | tail -4 | cut -c1-400
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_cuda_graph_gpu.py -m gpu -x -q 2>&1 | grep -v "UserWarning\|run_backward" | tail -6 | cut -c1-400
python tools/prof_cfg4_step.py 2>&1 | head -14
