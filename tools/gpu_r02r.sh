#!/bin/bash
# Round-2 pass R (1 GPU): GroupNorm with one CTA per (sample, group) — tests, kbench vs the cooperative fused kernel (variant 3) at B = 4 / 2 / 1
TAG=${1:-r02r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k groupnorm > $OUT/pytest_gn.log 2>&1; echo "pytest gn rc=$?"; tail -4 $OUT/pytest_gn.log | cut -c1-300
for b in 4 2; do
  for v in 0 3; do
    timeout 300 python tools/kbench.py --only groupnorm --batch $b --gn-variant $v > $OUT/kbench_gn_b${b}_v${v}.txt 2>&1; echo "kbench b$b v$v rc=$?"
    grep -E "^groupnorm" $OUT/kbench_gn_b${b}_v${v}.txt | cut -c1-110
  done
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/sanitize_memcheck.log
