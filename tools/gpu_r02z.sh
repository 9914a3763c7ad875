#!/bin/bash
# k1 with the TMA-store read wait moved off the step boundary: numerics, trace, kbench at B = 4 / 2 / 1
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tools/attn_check.py 0 > $OUT/attn_check.txt 2>&1; echo "attn_check rc=$?"; tail -1 $OUT/attn_check.txt
timeout 300 python tools/attn_check.py 22 > $OUT/attn_check22.txt 2>&1; echo "attn_check(22) rc=$?"; tail -1 $OUT/attn_check22.txt
TMX_LIB_PATH=$PWD/tweediemix_b200/lib/libtmx_trace.so python tools/attn_trace.py 1024 0 900 20 4 > $OUT/trace_n1024_b4.txt 2>&1
for b in 4 2 1; do
  timeout 300 python tools/kbench.py --only attention --batch $b --shapes Nk1024,Nk4096 > $OUT/kbench_attn_b$b.txt 2>&1; echo "kbench b$b rc=$?"
  grep -E "^attention" $OUT/kbench_attn_b$b.txt | cut -c1-110
done
