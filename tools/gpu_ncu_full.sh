#!/bin/bash
# ncu --set full captures of the hand-written kernels at the shapes of one fused K=3 step (1 GPU).
# Usage: bash tools/gpu_ncu_full.sh [tag]    -> gpurun_out/<tag>/*.ncu-rep ; summarise with tools/ncu_summary.py / ncu_stalls.py
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
# per shape kbench issues 3 warm-up launches + 2 graph replays (reps 1): skip 4 -> second replay
timeout 600 $NCU -k regex:attn_fwd -s 4 -c 1 -o $OUT/attn_n4096_final python tools/kbench.py --only attention --shapes Nq4096_Nk4096 --reps 1 > $OUT/ncu_attn_n4096.log 2>&1; echo "attn4096 rc=$?"
timeout 600 $NCU -k regex:attn_fwd -s 4 -c 1 -o $OUT/attn_n1024_final python tools/kbench.py --only attention --shapes Nq1024_Nk1024 --reps 1 > $OUT/ncu_attn_n1024.log 2>&1; echo "attn1024 rc=$?"
timeout 600 $NCU -k regex:attn_fwd -s 4 -c 1 -o $OUT/attn_cross_n4096 python tools/kbench.py --only attention --shapes Nq4096_Nk77 --reps 1 > $OUT/ncu_attn_cross.log 2>&1; echo "attn cross rc=$?"
# GroupNorm: ncu's kernel replay cannot re-run gn_fused_nhwc (inter-CTA barrier) -> profile the two-launch path
TMX_GN_TWO_PASS=1 timeout 600 $NCU -k regex:gn_ -s 8 -c 2 -o $OUT/gn_c320_hw16384 python tools/kbench.py --only groupnorm --shapes C320_HW16384 --reps 1 > $OUT/ncu_gn1.log 2>&1; echo "gn1 rc=$?"
TMX_GN_TWO_PASS=1 timeout 600 $NCU -k regex:gn_ -s 8 -c 2 -o $OUT/gn_c1280_hw1024 python tools/kbench.py --only groupnorm --shapes C1280_HW1024 --reps 1 > $OUT/ncu_gn2.log 2>&1; echo "gn2 rc=$?"
timeout 600 $NCU -k regex:blend_kernel -s 4 -c 1 -o $OUT/blend_2048 python tools/kbench.py --only blend --shapes imgs2048 --reps 1 > $OUT/ncu_blend.log 2>&1; echo "blend rc=$?"
timeout 600 $NCU -k regex:layernorm -s 4 -c 1 -o $OUT/ln_n1024 python tools/kbench.py --only layernorm --shapes N1024_D1280 --reps 1 > $OUT/ncu_ln.log 2>&1; echo "ln rc=$?"
timeout 600 $NCU -k regex:geglu -s 4 -c 1 -o $OUT/geglu_n1024 python tools/kbench.py --only geglu --reps 1 > $OUT/ncu_geglu.log 2>&1; echo "geglu rc=$?"
ls -la $OUT
