#!/bin/bash
# Round-2 pass I (1 GPU): video tests after the yardstick fix, the full kbench table with comparators (final), racecheck.
TAG=${1:-r02i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 600 python -m pytest tests/test_video.py -m gpu -q -s > $OUT/pytest_video.log 2>&1; echo "pytest video rc=$?"; tail -4 $OUT/pytest_video.log | cut -c1-300
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 $OUT/sanitize_racecheck.log
timeout 900 python tools/kbench.py --compare > $OUT/kbench.txt 2>&1; echo "kbench rc=$?"; grep -E "^video|^attention|per fused" $OUT/kbench.txt | cut -c1-200
