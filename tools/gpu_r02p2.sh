#!/bin/bash
# stream-K attention: where does the merge overhead come from?  timing-only variants (NOWAIT / NOMERGE give wrong results by design)
TAG=${1:-r02p2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for var in base nofence nowait nomerge; do
  lib=tweediemix_b200/lib/libtmx_$var.so
  [ $var = base ] && lib=tweediemix_b200/lib/libtmx.so
  for b in 2 1; do
    TMX_LIB_PATH=$PWD/$lib timeout 300 python tools/kbench.py --only attention --batch $b --shapes Nk1024,Nk4096 > $OUT/kb_${var}_b$b.txt 2>&1; echo "$var b$b rc=$?"
    grep -E "^attention" $OUT/kb_${var}_b$b.txt | cut -c1-100
  done
done
