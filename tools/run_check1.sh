# 1-GPU check: model + kernel tests, then the full default bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1s2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r1s2_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py ${BENCH_ARGS:---steps 30 --warmup 5} > gpurun_out/r1s2_n1.log 2>gpurun_out/r1s2_n1.err; echo "n1 rc=$?"
tail -3 gpurun_out/r1s2_n1.err | cut -c1-300
cut -c1-300 gpurun_out/r1s2_n1.log
