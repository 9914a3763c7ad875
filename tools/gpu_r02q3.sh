#!/bin/bash
TAG=${1:-r02q3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tools/attn_check.py 0 > $OUT/attn_check.txt 2>&1; echo "attn_check rc=$?"; tail -1 $OUT/attn_check.txt
for b in 4 2 1; do
  for v in 30 31 32; do
    timeout 300 python tools/kbench.py --only attention --batch $b --attn-variant $v --shapes Nk77 > $OUT/kbench_xattn_b${b}_v${v}.txt 2>&1; echo "kbench b$b v$v rc=$?"
    grep -E "^attention" $OUT/kbench_xattn_b${b}_v${v}.txt | cut -c1-110
  done
done
