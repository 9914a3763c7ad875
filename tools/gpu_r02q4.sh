#!/bin/bash
TAG=${1:-r02q4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for var in base x5 x6; do
  lib=tweediemix_b200/lib/libtmx_$var.so
  [ $var = base ] && lib=tweediemix_b200/lib/libtmx.so
  for b in 4 2; do
    TMX_LIB_PATH=$PWD/$lib timeout 300 python tools/kbench.py --only attention --batch $b --shapes Nk77 > $OUT/kb_${var}_b$b.txt 2>&1; echo "$var b$b rc=$?"
    grep -E "^attention" $OUT/kb_${var}_b$b.txt | cut -c1-100
  done
done
