#!/bin/bash
# k10 vs cuBLAS (+ epilogue kernel) at the GEMM shapes of a 2-row and a 1-row forward (2- / 4-rank concept-parallel groups)
TAG=${1:-r02u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for b in 2 1; do
  timeout 600 python tools/kbench.py --only linear --compare --batch $b > $OUT/kbench_linear_b$b.txt 2>&1; echo "kbench linear b$b rc=$?"
  grep -E "^linear |^cmp:cublas" $OUT/kbench_linear_b$b.txt | awk '{ if ($1 ~ /cmp:cublas-gemm/) printf "%-16s %-44s %8s us\n", $1, $5, $7; else if ($1 ~ /cmp/) printf "%-16s %-44s %8s us\n", $1, $7, $9; else printf "%-16s %-44s %8s us\n", $1, $2, $4 }'
done
