#!/bin/bash
# Round-2 pass B (1 GPU): k10 GEMM correctness + speed vs cuBLAS, full -m gpu suite (fused and library-GEMM paths), bench A/B.
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_linear.py -q > $OUT/pytest_linear.log 2>&1; echo "pytest linear rc=$?"; tail -25 $OUT/pytest_linear.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/sanitize_memcheck.log
timeout 600 python tools/kbench.py --only linear --compare > $OUT/kbench_linear.txt 2>&1; echo "kbench linear rc=$?"; grep -E "linear|cublas" $OUT/kbench_linear.txt
timeout 1800 python -m pytest tests -m gpu -q -s --ignore=tests/test_gpu_linear.py > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu (fused GEMM path) rc=$?"
grep -E "PARITY|passed|failed|FAILED|Error" $OUT/pytest_gpu.log | cut -c1-400 | tail -40
TMX_GEMM=cublas timeout 1800 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity.py -q -s > $OUT/pytest_gpu_cublas.log 2>&1; echo "pytest gpu (library GEMM path) rc=$?"
grep -E "PARITY|passed|failed|FAILED|Error" $OUT/pytest_gpu_cublas.log | cut -c1-400 | tail -30
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench (fused) rc=$?"; tail -3 $OUT/bench.err
TMX_GEMM=cublas timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $OUT/bench_cublas.json 2> $OUT/bench_cublas.err; echo "bench (cublas) rc=$?"; tail -3 $OUT/bench_cublas.err
timeout 900 python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_lora.json 2> $OUT/bench_lora.err; echo "bench lora (fused) rc=$?"; tail -3 $OUT/bench_lora.err
TMX_GEMM=cublas timeout 900 python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_lora_cublas.json 2> $OUT/bench_lora_cublas.err; echo "bench lora (cublas) rc=$?"; tail -3 $OUT/bench_lora_cublas.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02b/bench*.json")):
    try:
        d=json.load(open(f)); print(f, "value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), d["fused_step_tmx_kernel_ms"])
    except Exception as e: print(f, "unreadable", e)
PY
