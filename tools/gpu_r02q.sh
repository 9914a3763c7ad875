#!/bin/bash
# Round-2 pass Q (1 GPU): streaming short-K/V (cross-attention) kernel — numerics, kbench vs the tcgen05 path (variant 30), MT = 1 / 2
TAG=${1:-r02q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu -x > $OUT/pytest_attn.log 2>&1; echo "pytest attn rc=$?"; tail -5 $OUT/pytest_attn.log | cut -c1-300
for b in 4 2 1; do
  for v in 30 31 32; do
    timeout 300 python tools/kbench.py --only attention --batch $b --attn-variant $v --shapes Nk77 > $OUT/kbench_xattn_b${b}_v${v}.txt 2>&1; echo "kbench b$b v$v rc=$?"
    grep -E "^attention" $OUT/kbench_xattn_b${b}_v${v}.txt | cut -c1-110
  done
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/sanitize_memcheck.log
