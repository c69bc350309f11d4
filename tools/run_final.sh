# final numbers of a session: every -m gpu test, smoke(), the default bench line and the reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r2_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r2_bench_full.json 2>gpurun_out/r2_bench_full.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_full.err | cut -c1-300
cut -c1-300 gpurun_out/r2_bench_full.json
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2>gpurun_out/r2_bench_reference.err; echo "reference rc=$?"
cut -c1-400 gpurun_out/r2_bench_reference.json
