#!/bin/bash
for nv in 48 64 32; do echo "NV=$nv"; RBM_CE_WIDE_NV=$nv timeout 300 python tools/dbg_ce_wide.py time 2>&1 | grep "B=512" | tail -2; done
