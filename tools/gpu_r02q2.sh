#!/bin/bash
# ncu --set full of the short-K/V attention kernel (MT = 1) at the two SDXL cross-attention shapes
TAG=${1:-r02q2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:short_kv -s 4 -c 1 -o $OUT/xattn_n1024 python tools/kbench.py --only attention --shapes Nq1024_Nk77 --reps 1 > $OUT/ncu_x1024.log 2>&1; echo "x1024 rc=$?"
timeout 600 $NCU -k regex:short_kv -s 4 -c 1 -o $OUT/xattn_n4096 python tools/kbench.py --only attention --shapes Nq4096_Nk77 --reps 1 > $OUT/ncu_x4096.log 2>&1; echo "x4096 rc=$?"
ls -la $OUT
