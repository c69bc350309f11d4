# final numbers of the session: every -m gpu test, smoke(), the default bench line, then the ncu launch list of the same step
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1s2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r1s2_pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r1s2_bench_full.json 2>gpurun_out/r1s2_bench_full.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/r1s2_bench_full.json
if [ -n "$WITH_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1s2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_r1s2.log 2>&1; echo "ncu rc=$?"
fi
