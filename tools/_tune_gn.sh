#!/bin/bash
run() { TMX_NVCC_EXTRA="$1" python -m tweediemix_b200.build --force > /dev/null 2>&1; echo "== $1"; TMX_NVCC_EXTRA="$1" timeout 100 python tools/kbench.py --only groupnorm --shapes C1280_HW1024,C640_HW1024,C640_HW4096 2>&1 | grep -E "groupnorm "; }
run "-DTMX_GN_EXPERIMENT_STAGE=1"
run "-DTMX_GN_EXPERIMENT_STAGE=2"
run "-DTMX_GN_EXPERIMENT_STAGE=3"
python -m tweediemix_b200.build --force > /dev/null 2>&1
