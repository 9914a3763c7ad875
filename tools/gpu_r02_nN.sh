#!/bin/bash
# Multi-GPU bench pass: bash tools/gpu_r02_nN.sh N  (the default run: configs[1] + the extra legs that N GPUs allow)
N=${1:-4}
OUT=gpurun_out/${2:-r02n$N}
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt 2>&1
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 2 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench n$N rc=$?"; tail -4 $OUT/bench_n$N.err | cut -c1-300
python - $OUT/bench_n$N.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print("value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), d["scaling"], d["config"]["parallelism"], d["concept_parallel_check"])
    for o in d["other_configs"]:
        print(o.get("baseline_config"), o.get("error") or (round(o["value"],4), round(o["e2e"]["value"],4), round(o["per_denoise_step_ms"],2), o["config"]["parallelism"], o["concept_parallel_check"], o["sample_forwards_executed_per_image"]))
except Exception as e: print("unreadable", e)
PY
