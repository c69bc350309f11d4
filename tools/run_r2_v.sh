#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_models_gpu.py -m gpu -x -q 2>&1 | grep -v "UserWarning\|run_backward" | tail -8 | cut -c1-600
timeout 900 python bench.py --workload sasrec > gpurun_out/r2v_bench_sasrec.json 2> gpurun_out/r2v_bench_sasrec.err; echo "bench rc=$?"
tail -3 gpurun_out/r2v_bench_sasrec.err | cut -c1-300
cut -c1-600 gpurun_out/r2v_bench_sasrec.json
