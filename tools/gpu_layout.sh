#!/bin/bash
# k13 / k14 (channel concat, nearest 2x upsample): tests, kbench vs ATen, sanitizer, then the step A/B (old library = without them is not
# possible through TMX_LIB_PATH since the Python call sites changed: compare against profiles/r02final_bench.json of the same day instead)
OUT=gpurun_out/${1:-r02lay}; mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "cat_channels or upsample or layout" > $OUT/pytest_layout.log 2>&1; echo "pytest layout rc=$?"; tail -2 $OUT/pytest_layout.log | cut -c1-200
timeout 300 python tools/kbench.py --only layout --compare > $OUT/kbench_layout.txt 2>&1; echo "kbench rc=$?"; grep -E "^layout|^cmp:aten" $OUT/kbench_layout.txt | cut -c1-120
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -1 $OUT/sanitize_memcheck.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x > $OUT/pytest_model.log 2>&1; echo "pytest model rc=$?"; tail -2 $OUT/pytest_model.log | cut -c1-200
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), d["clocks"]["sm_mhz"], {k: round(v,2) for k,v in d["fused_step_tmx_kernel_ms"].items()})
PY
