#!/bin/bash
echo "default (ce_tc 3xTF32)"; python tools/prof_ce.py 2>&1 | tail -1
echo "ce_wide d=64"; RBM_CE_WIDE_D64=1 timeout 120 python tools/prof_ce.py 2>&1 | tail -2
RBM_CE_WIDE_D64=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "score_ce" 2>&1 | tail -4 | cut -c1-300
