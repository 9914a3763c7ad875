#!/bin/bash
# One GPU-box pass: parity tests, kernel micro-bench, the bench line, and the ncu launch list of the bench command.
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python -m tweediemix_b200.build --force > $OUT/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 600 python tools/kbench.py > $OUT/kbench.txt 2>&1; echo "kbench rc=$?"
cat $OUT/kbench.txt | tail -40
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
tail -c 6000 $OUT/bench.json
tail -5 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
cat $OUT/bench_ref.json
# launch list of one eager fused denoise step (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file $OUT/launches_fused_step.csv python bench.py --ncu-range --warmup 1 > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
wc -l $OUT/launches_fused_step.csv
