#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "topk" 2>&1 | grep -v "UserWarning\|run_backward" | tail -15 > gpurun_out/r2i_pytest_topk.log
cat gpurun_out/r2i_pytest_topk.log | cut -c1-800
timeout 600 python bench.py --workload eval --no-cpu-baseline --steps 5 > gpurun_out/r2i_bench_eval.json 2> gpurun_out/r2i_bench_eval.err
tail -c 300 gpurun_out/r2i_bench_eval.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2i_bench_eval.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d['metrics_last_step'])
"
RBM_TOPK_F16=0 timeout 600 python bench.py --workload eval --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('tf32 path:', d['value'], d['ms_per_step'])
"
