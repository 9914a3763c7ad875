#!/bin/bash
# Round-2 pass H (1 GPU): video-stage kernels (tests, sanitizer, kbench), full -m gpu regression, bench with the graph-timed roofline.
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 600 python -m pytest tests/test_video.py -m gpu -q -s > $OUT/pytest_video.log 2>&1; echo "pytest video rc=$?"; tail -6 $OUT/pytest_video.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/sanitize_memcheck.log
timeout 600 python tools/kbench.py --only video > $OUT/kbench_video.txt 2>&1; tail -6 $OUT/kbench_video.txt
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?"
python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), "roofline", round(d["roofline"]["frac"],3), d["roofline"]["timing"], {k:(round(v["avg_ms"]*1e3,1), round(v["tflops"])) for k,v in d["roofline"]["by_shape"].items()}, "cpu", d["cpu_baseline"]["value"])
PY
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.txt
