#!/bin/bash
# Round-2 pass L (1 GPU): resadd_ln at 4 CTAs/SM (kbench), new tests, bench.
TAG=${1:-r02l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 600 python tools/kbench.py --only resadd_ln,layernorm,resadd > $OUT/kbench_ln.txt 2>&1; cat $OUT/kbench_ln.txt | tail -9
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), {k: round(v,2) for k,v in d["fused_step_tmx_kernel_ms"].items()})
PY
