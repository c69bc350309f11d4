#!/bin/bash
# ncu launch list of the default bench step (after the same command exited 0 without ncu)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_list_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_list_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2_launches_bench_cfg2.csv
