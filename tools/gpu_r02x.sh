#!/bin/bash
# Round-2 pass X (1 GPU): ncu --set full of k2t at the two SDXL cross-attention shapes; compute-sanitizer racecheck + synccheck over the small-shape sweep
TAG=${1:-r02x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:short_kv_attn_tc -s 4 -c 1 -o $OUT/k2t_n1024 python tools/kbench.py --only attention --shapes Nq1024_Nk77 --reps 1 > $OUT/ncu_k2t_1024.log 2>&1; echo "k2t 1024 rc=$?"
timeout 600 $NCU -k regex:short_kv_attn_tc -s 4 -c 1 -o $OUT/k2t_n4096 python tools/kbench.py --only attention --shapes Nq4096_Nk77 --reps 1 > $OUT/ncu_k2t_4096.log 2>&1; echo "k2t 4096 rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 $OUT/sanitize_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -2 $OUT/sanitize_synccheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/sanitize_memcheck.log
