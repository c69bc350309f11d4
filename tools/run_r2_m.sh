#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dbg_ce_wide.py check 2>&1 | tail -8
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -x -q -k "score_ce or cfg4 or wide" 2>&1 | grep -v "UserWarning\|run_backward" | tail -5 | cut -c1-300
timeout 300 python tools/dbg_ce_wide.py time 2>&1 | tail -6
