#!/bin/bash
# the other two quantities of the metric as top-level bench lines (own arm, then the CPU arm on the reference's classes)
mkdir -p gpurun_out
for w in sasrec eval; do
  timeout 900 python bench.py --workload $w --no-extras > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "$w rc=$?"
  cut -c1-260 gpurun_out/r2_bench_$w.json
  timeout 900 python bench.py --impl reference --workload $w > gpurun_out/r2_bench_reference_$w.json 2> gpurun_out/r2_bench_reference_$w.err; echo "reference $w rc=$?"
  cut -c1-260 gpurun_out/r2_bench_reference_$w.json
done
