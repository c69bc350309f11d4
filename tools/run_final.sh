# final round numbers: the default bench line, then the ncu launch list of the same step (cold-cache, serialised: compare SHARES)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r1s2_bench_full.json 2>gpurun_out/r1s2_bench_full.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/r1s2_bench_full.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r1s2_bench_short.log 2>&1; echo "short rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_r2.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/launches_r2.csv
