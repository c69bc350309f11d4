#!/bin/bash
mkdir -p gpurun_out
for w in bert sasrec; do
timeout 900 python bench.py --workload $w --no-extras --no-cpu-baseline > gpurun_out/r2y_$w.json 2> gpurun_out/r2y_$w.err; echo "bench $w rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2y_$w.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
PY
done
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "graph" 2>&1 | tail -2
