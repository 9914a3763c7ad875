#!/bin/bash
# Round-2 pass T (1 GPU): ncu --set full of the final kernels — stream-K attention (N = 4096 / 1024), the streaming cross-attention kernel,
# the slab GroupNorm (single CTA per group at C=1280 32x32; C=640 64x64; 2-CTA cluster at C=1280 64x64)
TAG=${1:-r02t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:attn_fwd -s 4 -c 1 -o $OUT/attn_n4096_streamk python tools/kbench.py --only attention --shapes Nq4096_Nk4096 --reps 1 > $OUT/ncu_a4096.log 2>&1; echo "attn4096 rc=$?"
timeout 600 $NCU -k regex:attn_fwd -s 4 -c 1 -o $OUT/attn_n1024_streamk python tools/kbench.py --only attention --shapes Nq1024_Nk1024 --reps 1 > $OUT/ncu_a1024.log 2>&1; echo "attn1024 rc=$?"
timeout 600 $NCU -k regex:short_kv -s 4 -c 1 -o $OUT/xattn_n1024_final python tools/kbench.py --only attention --shapes Nq1024_Nk77 --reps 1 > $OUT/ncu_x1024.log 2>&1; echo "x1024 rc=$?"
timeout 600 $NCU -k regex:gn_group_slab -s 4 -c 1 -o $OUT/gn_slab_c1280_hw1024 python tools/kbench.py --only groupnorm --shapes C1280_HW1024 --reps 1 > $OUT/ncu_g1.log 2>&1; echo "gn1 rc=$?"
timeout 600 $NCU -k regex:gn_group_slab -s 4 -c 1 -o $OUT/gn_slab_c640_hw4096 python tools/kbench.py --only groupnorm --shapes C640_HW4096 --reps 1 > $OUT/ncu_g2.log 2>&1; echo "gn2 rc=$?"
ls -la $OUT | head -20
