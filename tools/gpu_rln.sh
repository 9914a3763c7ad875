#!/bin/bash
OUT=gpurun_out/${1:-r02rln}; mkdir -p $OUT
for var in base rln4 rln2; do
  lib=tweediemix_b200/lib/libtmx_$var.so; [ $var = base ] && lib=tweediemix_b200/lib/libtmx.so
  for b in 4 2; do TMX_LIB_PATH=$PWD/$lib timeout 300 python tools/kbench.py --only resadd_ln --batch $b > $OUT/kb_${var}_b$b.txt 2>&1; echo "$var b$b rc=$?"; grep -E "^resadd_ln" $OUT/kb_${var}_b$b.txt | cut -c1-100; done
done
