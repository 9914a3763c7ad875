#!/bin/bash
# Round-2 pass D (1 GPU): k10 after the epilogue restructure (16-column passes, coalesced split-K workspace, A&S erf).
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_linear.py -q > $OUT/pytest_linear.log 2>&1; echo "pytest linear rc=$?"; tail -8 $OUT/pytest_linear.log | cut -c1-300
timeout 600 python tools/kbench.py --only linear --compare > $OUT/kbench_linear.txt 2>&1; echo "kbench linear rc=$?"
grep -E "^linear|^cmp:cublas" $OUT/kbench_linear.txt | awk '{ if ($1 ~ /cmp:cublas-gemm/) printf "%-16s %-44s %8s us\n", $1, $5, $7; else if ($1 ~ /cmp/) printf "%-16s %-44s %8s us\n", $1, $7, $9; else printf "%-16s %-44s %8s us\n", $1, $2, $4 }'
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench (fused) rc=$?"; tail -3 $OUT/bench.err
timeout 900 python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_lora.json 2> $OUT/bench_lora.err; echo "bench lora (fused) rc=$?"; tail -3 $OUT/bench_lora.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/$TAG/bench*.json")):
    try:
        d=json.load(open(f)); print(f, "value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), {k: round(v,2) for k,v in d["fused_step_tmx_kernel_ms"].items()})
    except Exception as e: print(f, "unreadable", e)
PY
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:linear_kernel -s 4 -c 1 -o $OUT/linear_out_d1280 python tools/kbench.py --only linear --shapes "out|q_M4096" --reps 1 > $OUT/ncu_lin2.log 2>&1; echo "ncu out rc=$?"
