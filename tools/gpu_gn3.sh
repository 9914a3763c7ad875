#!/bin/bash
OUT=gpurun_out/${1:-r02gn4}; mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k groupnorm > $OUT/pytest_gn.log 2>&1; echo "pytest gn rc=$?"; tail -2 $OUT/pytest_gn.log | cut -c1-200
for b in 4 2; do timeout 300 python tools/kbench.py --only groupnorm --batch $b > $OUT/kb_b$b.txt 2>&1; echo "kbench b$b rc=$?"; grep -E "^groupnorm" $OUT/kb_b$b.txt | cut -c1-100; done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -1 $OUT/sanitize_memcheck.log
