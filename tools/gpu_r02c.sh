#!/bin/bash
# Round-2 pass C (1 GPU): k10 with the split-K tail — tests, sanitizer, kbench vs cuBLAS, bench A/B, ncu evidence.
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_linear.py -q > $OUT/pytest_linear.log 2>&1; echo "pytest linear rc=$?"; tail -15 $OUT/pytest_linear.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/sanitize_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/sanitize_racecheck.log
timeout 600 python tools/kbench.py --only linear --compare > $OUT/kbench_linear.txt 2>&1; echo "kbench linear rc=$?"
grep -E "^linear|^cmp:cublas" $OUT/kbench_linear.txt | awk '{printf "%-16s %-58s %8s us %8s TF\n", $1, ($1 ~ /cmp/ ? $2" "$3" "$4" "$5" "$6" "$7 : $2), ($1 ~ /cmp:cublas-gemm/ ? $7 : ($1 ~ /cmp/ ? $9 : $4)), ""}' | head -80
timeout 1800 python -m pytest tests -m gpu -q -s --ignore=tests/test_gpu_linear.py > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
grep -E "PARITY|passed|failed|FAILED|Error" $OUT/pytest_gpu.log | cut -c1-300 | tail -20
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench (fused) rc=$?"; tail -3 $OUT/bench.err
TMX_GEMM=cublas timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $OUT/bench_cublas.json 2> $OUT/bench_cublas.err; echo "bench (cublas) rc=$?"; tail -3 $OUT/bench_cublas.err
timeout 900 python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_lora.json 2> $OUT/bench_lora.err; echo "bench lora (fused) rc=$?"; tail -3 $OUT/bench_lora.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02c/bench*.json")):
    try:
        d=json.load(open(f)); print(f, "value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), {k: round(v,2) for k,v in d["fused_step_tmx_kernel_ms"].items()})
    except Exception as e: print(f, "unreadable", e)
PY
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:linear_kernel -s 4 -c 1 -o $OUT/linear_ff1_d1280 python tools/kbench.py --only linear --shapes ff1_M4096 --reps 1 > $OUT/ncu_lin1.log 2>&1; echo "ncu ff1 rc=$?"
timeout 600 $NCU -k regex:linear_kernel -s 4 -c 1 -o $OUT/linear_out_d1280 python tools/kbench.py --only linear --shapes "out|q_M4096" --reps 1 > $OUT/ncu_lin2.log 2>&1; echo "ncu out rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file $OUT/launches_fused_step.csv python bench.py --ncu-range --warmup 1 > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
wc -l $OUT/launches_fused_step.csv
