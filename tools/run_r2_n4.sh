#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 4 \
    > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
grep -v "UserWarning\|run_backward\|OMP_NUM\|\*\*\*\*" gpurun_out/r2_bench_n4.err | tail -10
grep "^{" gpurun_out/r2_bench_n4.json | head -c 300
