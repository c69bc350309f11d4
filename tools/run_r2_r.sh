#!/bin/bash
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "last_position" 2>&1 | grep -v "UserWarning\|run_backward" | tail -15 | cut -c1-500
for nv in 48 64 32; do echo "NV=$nv"; RBM_CE_WIDE_NV=$nv timeout 300 python tools/dbg_ce_wide.py time 2>&1 | grep "B=512" | tail -2; done
