#!/bin/bash
# A/B on ONE box: bench.py with the library of an earlier commit (TMX_LIB_PATH) and with the current one, interleaved (box-to-box clock /
# power differences are +-3 %, larger than most single-kernel gains).  Usage: bash tools/gpu_ab.sh <tag> <old-lib-name> [bench args]
TAG=${1:-ab}; OLD=${2:-libtmx_r02o.so}; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
for rep in 1 2; do
  for which in old new; do
    if [ $which = old ]; then export TMX_LIB_PATH=$PWD/tweediemix_b200/lib/$OLD; else unset TMX_LIB_PATH; fi
    timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $OUT/bench_${which}_$rep.json 2> $OUT/bench_${which}_$rep.err; echo "bench $which $rep rc=$?"
  done
done
unset TMX_LIB_PATH
python - $OUT <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], "value", round(d["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), "mhz", d["clocks"]["sm_mhz"], "roof", round(d["roofline"]["frac"],3), {k: round(v,2) for k,v in d["fused_step_tmx_kernel_ms"].items()})
PY
