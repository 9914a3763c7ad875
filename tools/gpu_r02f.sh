#!/bin/bash
# Round-2 pass F (1 GPU): GEMM policy 'auto' — full -m gpu suite, bench under the three policies, kbench (all families, comparators),
# launch list of one fused step, ncu of the GEGLU GEMM.
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
grep -E "PARITY|sampler parity|adapter|passed|failed|FAILED|Error" $OUT/pytest_gpu.log | cut -c1-260 | tail -40
for impl in auto tmx cublas; do
  TMX_GEMM=$impl timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $OUT/bench_$impl.json 2> $OUT/bench_$impl.err; echo "bench $impl rc=$?"; tail -2 $OUT/bench_$impl.err
done
for impl in auto cublas; do
  TMX_GEMM=$impl timeout 900 python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_lora_$impl.json 2> $OUT/bench_lora_$impl.err; echo "bench lora $impl rc=$?"; tail -2 $OUT/bench_lora_$impl.err
done
python - "$OUT" <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1] + "/bench*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), {k: round(v,2) for k,v in d["fused_step_tmx_kernel_ms"].items()})
    except Exception as e: print(f, "unreadable", e)
PY
timeout 900 python tools/kbench.py --compare > $OUT/kbench.txt 2>&1; echo "kbench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file $OUT/launches_fused_step.csv python bench.py --ncu-range --warmup 1 > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:linear_kernel -s 4 -c 1 -o $OUT/linear_ff1_d1280 python tools/kbench.py --only linear --shapes ff1_M4096 --reps 1 > $OUT/ncu_lin1.log 2>&1; echo "ncu ff1 rc=$?"
