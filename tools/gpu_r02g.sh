#!/bin/bash
# Round-2 pass G (1 GPU): attention with the TMA-store epilogue — tests, sanitizer, kbench, bench.
TAG=${1:-r02g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_attention.py -q > $OUT/pytest_attn.log 2>&1; echo "pytest attention rc=$?"; tail -5 $OUT/pytest_attn.log | cut -c1-300
for v in 0 12; do timeout 300 python tools/attn_check.py $v > $OUT/attn_check_$v.txt 2>&1; echo "attn_check $v rc=$?"; tail -1 $OUT/attn_check_$v.txt; done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/sanitize_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/sanitize_racecheck.log
timeout 600 python tools/kbench.py --only attention > $OUT/kbench_attn.txt 2>&1; cat $OUT/kbench_attn.txt | tail -6
timeout 2400 python -m pytest tests -m gpu -q -x --ignore=tests/test_gpu_linear.py --ignore=tests/test_gpu_attention.py > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), "roofline", round(d["roofline"]["frac"],3), {k:(round(v["avg_ms"]*1e3,1), round(v["tflops"])) for k,v in d["roofline"]["by_shape"].items()})
PY
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:attn_fwd -s 4 -c 1 -o $OUT/attn_n4096_final python tools/kbench.py --only attention --shapes Nq4096_Nk4096 --reps 1 > $OUT/ncu_attn_n4096.log 2>&1; echo "attn4096 rc=$?"
timeout 600 $NCU -k regex:attn_fwd -s 4 -c 1 -o $OUT/attn_n1024_final python tools/kbench.py --only attention --shapes Nq1024_Nk1024 --reps 1 > $OUT/ncu_attn_n1024.log 2>&1; echo "attn1024 rc=$?"
timeout 600 $NCU -k regex:attn_fwd -s 4 -c 1 -o $OUT/attn_cross_n1024 python tools/kbench.py --only attention --shapes Nq1024_Nk77 --reps 1 > $OUT/ncu_attn_cross.log 2>&1; echo "attn cross rc=$?"
