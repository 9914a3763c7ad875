#!/bin/bash
# Round-2 pass J (1 GPU): video tests, k10 with the eager split-K mode (+2000) per shape.
TAG=${1:-r02j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 600 python -m pytest tests/test_video.py -m gpu -q -s > $OUT/pytest_video.log 2>&1; echo "pytest video rc=$?"; tail -3 $OUT/pytest_video.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_linear.py -q -k split > $OUT/pytest_linear_split.log 2>&1; echo "pytest linear split rc=$?"; tail -3 $OUT/pytest_linear_split.log | cut -c1-300
timeout 600 python tools/kbench.py --only linear --compare > $OUT/kbench_linear.txt 2>&1; echo "kbench linear rc=$?"
grep -E "^linear|^cmp:cublas" $OUT/kbench_linear.txt | awk '{ if ($1 ~ /cmp:cublas-gemm/) printf "%-16s %-44s %8s us\n", $1, $5, $7; else if ($1 ~ /cmp/) printf "%-16s %-44s %8s us\n", $1, $7, $9; else printf "%-16s %-44s %8s us\n", $1, $2, $4 }'
