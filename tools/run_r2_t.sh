#!/bin/bash
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "rows or sas" 2>&1 | grep -v "UserWarning\|run_backward" | tail -25 | cut -c1-600
python tools/prof_sas_step.py 2>&1 | tail -26
