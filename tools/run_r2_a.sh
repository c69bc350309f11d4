#!/bin/bash
# round 2, GPU call A: full GPU test suite + the default bench line (1 GPU)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.log
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 3000 gpurun_out/r2a_pytest.log
tail -c 1500 gpurun_out/r2a_bench.err
head -c 6000 gpurun_out/r2a_bench.json
