#!/bin/bash
# 2-GPU pass: the NCCL parity test (tests/test_gpu_nccl.py) and the bench at N = 2 (one concept-parallel group of 2 ranks).
OUT=gpurun_out/${1:-r02n2}
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt 2>&1
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_nccl.py -m gpu -q -s > $OUT/pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"; grep -E "NCCL parity|passed|failed|skipped|Error" $OUT/pytest_nccl.log | cut -c1-600 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 2 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?"; tail -4 $OUT/bench_n2.err | cut -c1-300
python - $OUT/bench_n2.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print("N=2 value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "step_ms", round(d["per_denoise_step_ms"],2), d["scaling"], d["config"]["parallelism"], d["concept_parallel_check"], "roof", round(d["roofline"]["frac"],3), {k: round(v["avg_ms"]*1e3,1) for k,v in d["roofline"]["by_shape"].items()})
except Exception as e: print("unreadable", e)
PY
