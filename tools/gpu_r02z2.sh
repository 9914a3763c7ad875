#!/bin/bash
# k1: P V(n-1) wait behind the exp block (default) vs in front of it (-DTMX_ATTN_PV_WAIT_EARLY), same box
TAG=${1:-r02z2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tools/attn_check.py 0 > $OUT/attn_check.txt 2>&1; echo "attn_check rc=$?"; tail -1 $OUT/attn_check.txt
timeout 300 python tools/attn_check.py 22 > $OUT/attn_check22.txt 2>&1; echo "attn_check(22) rc=$?"; tail -1 $OUT/attn_check22.txt
for rep in 1 2; do for var in early late; do
  if [ $var = early ]; then export TMX_LIB_PATH=$PWD/tweediemix_b200/lib/libtmx_early.so; else unset TMX_LIB_PATH; fi
  timeout 300 python tools/kbench.py --only attention --batch 4 --shapes Nk1024,Nk4096 > $OUT/kbench_${var}_$rep.txt 2>&1; echo "kbench $var $rep rc=$?"
  grep -E "^attention" $OUT/kbench_${var}_$rep.txt | cut -c1-110
done; done
unset TMX_LIB_PATH
timeout 900 python -m pytest tests/test_gpu_attention.py -q -m gpu -x > $OUT/pytest_attn.log 2>&1; echo "pytest attn rc=$?"; tail -2 $OUT/pytest_attn.log | cut -c1-200
