#!/bin/bash
# round 2, GPU call E (1 GPU): new CE parity tests, default bench, ncu of the wide CE kernels
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -x -q -k "score_ce or cfg4 or wide" 2>&1 | grep -v "UserWarning\|run_backward" | tail -15 > gpurun_out/r2e_pytest.log
cat gpurun_out/r2e_pytest.log
python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -c 600 gpurun_out/r2e_bench.err
B=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ce_wide_kernel -c 3 -f -o gpurun_out/prof_ce_wide_r2 python tools/dbg_ce_wide.py prof > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_ncu.log
