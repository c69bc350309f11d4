#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -x -q -k "score_ce or cfg4 or wide" 2>&1 | grep -v "UserWarning\|run_backward" | tail -8 > gpurun_out/r2f_pytest.log
cat gpurun_out/r2f_pytest.log | cut -c1-400
python tools/prof_cfg4_step.py > gpurun_out/r2f_cfg4_prof.log 2>&1
tail -30 gpurun_out/r2f_cfg4_prof.log
