#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_dist_gpu.py -m gpu -x -q -rs 2>&1 | tail -15 > gpurun_out/r2b_dist_pytest.log
tail -4 gpurun_out/r2b_dist_pytest.log | cut -c1-300
timeout 1200 python -m pytest tests/test_models_gpu.py -m gpu -x -q 2>&1 | tail -3 | cut -c1-300
