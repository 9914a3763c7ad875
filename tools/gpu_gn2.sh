#!/bin/bash
# two-launch GroupNorm (variant 1) at the sites that take it: register cap / rows in flight variants, same box
OUT=gpurun_out/${1:-r02gn2}; mkdir -p $OUT
for var in gn_b1u8 gn_b1u8s16 gn_b1u6s12; do
  lib=tweediemix_b200/lib/libtmx_$var.so; [ $var = base ] && lib=tweediemix_b200/lib/libtmx.so
  TMX_LIB_PATH=$PWD/$lib timeout 300 python tools/kbench.py --only groupnorm --gn-variant 1 --shapes HW16384,C960_HW4096,C1920_HW4096 > $OUT/kb_$var.txt 2>&1; echo "$var rc=$?"
  grep -E "^groupnorm" $OUT/kb_$var.txt | cut -c1-100
done
