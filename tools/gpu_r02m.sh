#!/bin/bash
# Round-2 pass M (1 GPU): k10 with 320-wide single-accumulator tiles — tests, sanitizer, kbench vs cuBLAS, bench (auto / tmx).
TAG=${1:-r02m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_linear.py -q > $OUT/pytest_linear.log 2>&1; echo "pytest linear rc=$?"; tail -6 $OUT/pytest_linear.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/sanitize_memcheck.log
timeout 600 python tools/kbench.py --only linear --compare > $OUT/kbench_linear.txt 2>&1; echo "kbench linear rc=$?"
grep -E "^linear|^cmp:cublas" $OUT/kbench_linear.txt | awk '{ if ($1 ~ /cmp:cublas-gemm/) printf "%-16s %-44s %8s us\n", $1, $5, $7; else if ($1 ~ /cmp/) printf "%-16s %-44s %8s us\n", $1, $7, $9; else printf "%-16s %-44s %8s us\n", $1, $2, $4 }'
for impl in auto tmx; do
  TMX_GEMM=$impl timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $OUT/bench_$impl.json 2> $OUT/bench_$impl.err; echo "bench $impl rc=$?"; tail -2 $OUT/bench_$impl.err
done
python - $OUT <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1] + "/bench*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], "value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), {k: round(v,2) for k,v in d["fused_step_tmx_kernel_ms"].items()})
PY
