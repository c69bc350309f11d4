#!/bin/bash
# round 2 evidence: ncu launch list of the default bench (shares), ncu --set full of the new kernels
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2p_bench_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2p_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_pair" -s 4 -c 2 -f -o gpurun_out/prof_attn_pair_r2 \
    python tools/prof_sas_step.py > gpurun_out/r2p_ncu_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_topk_kernel" -s 2 -c 1 -f -o gpurun_out/prof_topk_r2 \
    python tools/prof_topk.py > gpurun_out/r2p_ncu_topk.log 2>&1
B=128 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dw16_kernel|tc_linear_persistent" -s 40 -c 6 -f -o gpurun_out/prof_linear_wide_r2 \
    python tools/prof_cfg4_step.py > gpurun_out/r2p_ncu_lin.log 2>&1
tail -2 gpurun_out/r2p_ncu_pair.log gpurun_out/r2p_ncu_topk.log gpurun_out/r2p_ncu_lin.log
