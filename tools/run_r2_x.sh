#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | grep -v "UserWarning\|run_backward" | tail -12 | cut -c1-600
timeout 1200 python -m pytest tests/test_models_gpu.py -m gpu -x -q 2>&1 | grep -v "UserWarning\|run_backward" | tail -12 | cut -c1-800
timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2w_bench_bert.json 2> gpurun_out/r2w_bench_bert.err; echo "bench rc=$?"
tail -3 gpurun_out/r2w_bench_bert.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench_bert.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_time_shares'])
PY
