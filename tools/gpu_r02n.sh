#!/bin/bash
# Round-2 pass N (1 GPU): GroupNorm fused kernel — cooperative launch vs plain launch of the same kernel.
OUT=gpurun_out/r02n
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 300 python tools/kbench.py --only groupnorm > $OUT/kbench_gn_coop.txt 2>&1; tail -14 $OUT/kbench_gn_coop.txt
timeout 300 python tools/kbench.py --only groupnorm --gn-variant 2 > $OUT/kbench_gn_plain.txt 2>&1; tail -14 $OUT/kbench_gn_plain.txt
