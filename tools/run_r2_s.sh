#!/bin/bash
echo "deep=0"; RBM_LINEAR_DEEP=0 python tools/time_linear_cfg2.py 2>&1 | tail -4
echo "deep=1"; RBM_LINEAR_DEEP=1 python tools/time_linear_cfg2.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "linear" 2>&1 | tail -2
