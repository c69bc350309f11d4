set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dist_gpu.py -x -q > gpurun_out/r1s2_dist_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r1s2_dist_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus 2 --steps 30 --warmup 5 --step-mode graph --dp-collective split > gpurun_out/r1s2_dp2_split.log 2>gpurun_out/r1s2_dp2_split.err; echo "split rc=$?"
tail -c 1500 gpurun_out/r1s2_dp2_split.log | cut -c1-700
timeout 150 $TR bench.py --gpus 2 --steps 30 --warmup 5 --step-mode graph --dp-collective graph > gpurun_out/r1s2_dp2_graph.log 2>gpurun_out/r1s2_dp2_graph.err; echo "graph rc=$?"
tail -c 1500 gpurun_out/r1s2_dp2_graph.log | cut -c1-400
tail -5 gpurun_out/r1s2_dp2_graph.err
timeout 200 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r1s2_n1_graph.log 2>gpurun_out/r1s2_n1_graph.err; echo "n1 rc=$?"
cut -c1-700 gpurun_out/r1s2_n1_graph.log
