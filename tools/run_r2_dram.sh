#!/bin/bash
# DRAM bytes of every launch of the default bench step (ncu, cold-cache per launch: an upper bound on the bytes of the replayed graph)
mkdir -p gpurun_out
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_dram_bench_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_dram_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2_dram_bench_cfg2.csv
