#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | grep -v "UserWarning\|run_backward" | tail -25 > gpurun_out/r2h_pytest_attn.log
cat gpurun_out/r2h_pytest_attn.log | cut -c1-600
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "sas" 2>&1 | grep -v "UserWarning\|run_backward" | tail -12 > gpurun_out/r2h_pytest_sas.log
cat gpurun_out/r2h_pytest_sas.log | cut -c1-600
timeout 300 python bench.py --workload sasrec --no-cpu-baseline --no-extras --steps 30 > gpurun_out/r2h_bench_sas.json 2> gpurun_out/r2h_bench_sas.err
tail -c 300 gpurun_out/r2h_bench_sas.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2h_bench_sas.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['kernel_time_shares'], d['roofline'])
"
timeout 200 python tools/dbg_ce_wide.py time 2>&1 | tail -6
