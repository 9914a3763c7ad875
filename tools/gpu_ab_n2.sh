#!/bin/bash
# A/B at N = 2 on ONE box (one concept-parallel group of two ranks): library of an earlier commit vs the current one, interleaved.
TAG=${1:-ab_n2}; OLD=${2:-libtmx_r02o.so}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m tweediemix_b200.build > $OUT/build.log 2>&1
for rep in 1 2; do
  for which in old new; do
    if [ $which = old ]; then export TMX_LIB_PATH=$PWD/tweediemix_b200/lib/$OLD; else unset TMX_LIB_PATH; fi
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_${which}_$rep.json 2> $OUT/bench_${which}_$rep.err; echo "bench $which $rep rc=$?"
  done
done
unset TMX_LIB_PATH
python - $OUT <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], "value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), "mhz", d["clocks"]["sm_mhz"], "roof", round(d["roofline"]["frac"],3), {k:(round(v["avg_ms"]*1e3,1)) for k,v in d["roofline"]["by_shape"].items()}, d["concept_parallel_check"]["rel"])
PY
