#!/bin/bash
# round 2, GPU call B (2 GPUs): the 2-GPU NCCL test at HEAD + bench.py --gpus 2 (DP headline + the sharded extras)
mkdir -p gpurun_out
git_rev=$(cat .git_rev 2>/dev/null)
python -m pytest tests/test_dist_gpu.py -m gpu -x -q -rs 2>&1 | tail -15 > gpurun_out/r2b_dist_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 \
    > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
cat gpurun_out/r2b_dist_pytest.log
tail -c 2500 gpurun_out/r2b_bench_n2.err
head -c 300 gpurun_out/r2b_bench_n2.json
