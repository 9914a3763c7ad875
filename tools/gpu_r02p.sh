#!/bin/bash
# Round-2 pass P (1 GPU): attention with the stream-K schedule (units cut along K/V between CTAs) — numerics with the split forced on,
# kbench at B = 4 / 2 / 1 under the three schedules (20 = whole tiles, 22 = always split, 0 = cost model), attention tests.
TAG=${1:-r02p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 300 python tools/attn_check.py 22 > $OUT/attn_check_split.txt 2>&1; echo "attn_check(22) rc=$?"; tail -2 $OUT/attn_check_split.txt
timeout 300 python tools/attn_check.py 0 > $OUT/attn_check.txt 2>&1; echo "attn_check(0) rc=$?"; tail -1 $OUT/attn_check.txt
for b in 4 2 1; do
  for v in 20 22 0; do
    timeout 300 python tools/kbench.py --only attention --batch $b --attn-variant $v --shapes Nk1024,Nk4096 > $OUT/kbench_attn_b${b}_v${v}.txt 2>&1; echo "kbench b$b v$v rc=$?"
    grep -E "^attention" $OUT/kbench_attn_b${b}_v${v}.txt | cut -c1-110
  done
done
timeout 900 python -m pytest tests/test_gpu_attention.py -q -m gpu > $OUT/pytest_attn.log 2>&1; echo "pytest attn rc=$?"; tail -3 $OUT/pytest_attn.log | cut -c1-300
