#!/bin/bash
# round 2, GPU call G (8 GPUs): bench.py --gpus 8 exactly as the driver launches it
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 \
    > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err
grep -v "UserWarning\|run_backward\|OMP_NUM\|\*\*\*\*" gpurun_out/r2g_bench_n8.err | tail -20
grep "^{" gpurun_out/r2g_bench_n8.json | head -c 400
