#!/bin/bash
# ncu launch list (gpu__time_duration.sum per launch, --clock-control none) of ONE eager fused K=3 step with the current kernels
TAG=${1:-r02final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file $OUT/launches_fused_step.csv python bench.py --ncu-range --warmup 1 > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
wc -l $OUT/launches_fused_step.csv
python tools/launch_summary.py $OUT/launches_fused_step.csv > $OUT/launches_fused_step.summary.txt; head -30 $OUT/launches_fused_step.summary.txt | cut -c1-150
