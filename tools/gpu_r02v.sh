#!/bin/bash
# Round-2 pass V (1 GPU): tcgen05 short-K/V attention kernel (k2t) — numerics under variant 33, kbench vs k2s (31) and k1 (30)
TAG=${1:-r02v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 300 python tools/attn_check.py 33 > $OUT/attn_check_k2t.txt 2>&1; echo "attn_check(33) rc=$?"; grep -E "Nk77|Nk65|Nk80|ok|Error|error" $OUT/attn_check_k2t.txt | tail -14
for b in 4 2; do
  for v in 31 33; do
    timeout 300 python tools/kbench.py --only attention --batch $b --attn-variant $v --shapes Nk77 > $OUT/kbench_xattn_b${b}_v${v}.txt 2>&1; echo "kbench b$b v$v rc=$?"
    grep -E "^attention" $OUT/kbench_xattn_b${b}_v${v}.txt | cut -c1-110
  done
done
