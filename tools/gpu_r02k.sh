#!/bin/bash
# Round-2 pass K (1 GPU): two-source GroupNorm (cat-free up blocks) — tests, sanitizer, full regression, bench A/B.
TAG=${1:-r02k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "groupnorm" > $OUT/pytest_gn.log 2>&1; echo "pytest gn rc=$?"; tail -4 $OUT/pytest_gn.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/sanitize_memcheck.log
timeout 2400 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_linear.py > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 $OUT/pytest_gpu.log | cut -c1-300
TMX_CAT_FREE=0 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $OUT/bench_cat.json 2> $OUT/bench_cat.err; echo "bench (torch.cat) rc=$?"
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench (cat-free) rc=$?"; tail -2 $OUT/bench.err
python - $OUT <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1] + "/bench*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], "value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), {k: round(v,2) for k,v in d["fused_step_tmx_kernel_ms"].items()})
PY
