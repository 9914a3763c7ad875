#!/bin/bash
# Final pass (1 GPU): what the driver runs at round end — build, pytest -m gpu, smoke(), bench.py (both arms) — plus the full kbench table.
TAG=${1:-r02final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1; echo "build rc=$?"
timeout 2400 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.txt
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?"
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), "roofline", round(d["roofline"]["frac"],3), {k:(round(v["avg_ms"]*1e3,1), round(v["tflops"])) for k,v in d["roofline"]["by_shape"].items()}, [ (o["kernel"][:12], round(o.get("frac",0),2)) for o in d["roofline_other"]], "cpu", d["cpu_baseline"]["value"], d["clocks"])
print({k: round(v,2) for k,v in d["fused_step_tmx_kernel_ms"].items()})
PY
timeout 900 python tools/kbench.py > $OUT/kbench.txt 2>&1; echo "kbench rc=$?"; tail -2 $OUT/kbench.txt | cut -c1-300
