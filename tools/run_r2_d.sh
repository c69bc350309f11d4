#!/bin/bash
# wide CE bring-up: parity check + timing (1 GPU)
mkdir -p gpurun_out
timeout 300 python tools/dbg_ce_wide.py check > gpurun_out/r2d_ce_check.log 2>&1
tail -12 gpurun_out/r2d_ce_check.log
timeout 300 python tools/dbg_ce_wide.py time > gpurun_out/r2d_ce_time.log 2>&1
tail -8 gpurun_out/r2d_ce_time.log
RBM_CE_WIDE_PAIR=0 timeout 300 python tools/dbg_ce_wide.py time > gpurun_out/r2d_ce_time_nopair.log 2>&1
tail -8 gpurun_out/r2d_ce_time_nopair.log
