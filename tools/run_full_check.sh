# full GPU check: every -m gpu test (2 GPUs: incl. the NCCL tests), then the bench at N=1 and N=2
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1s2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r1s2_pytest_gpu.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r1s2_n1.log 2>gpurun_out/r1s2_n1.err; echo "n1 rc=$?"
cut -c1-400 gpurun_out/r1s2_n1.log
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r1s2_n2.log 2>gpurun_out/r1s2_n2.err; echo "n2 rc=$?"
tail -1 gpurun_out/r1s2_n2.log | cut -c1-300
fi
