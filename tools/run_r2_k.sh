#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "linear" 2>&1 | grep -v "UserWarning\|run_backward" | tail -30 > gpurun_out/r2k_pytest_linear.log
cat gpurun_out/r2k_pytest_linear.log | cut -c1-500
timeout 300 python bench.py --workload sasrec --no-cpu-baseline --no-extras --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('sasrec', d['value'], d['ms_per_step'], d['kernel_time_shares'])
"
