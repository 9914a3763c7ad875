#!/bin/bash
# Bottleneck-location experiments for the attention kernel (GPU box): results of the EXPERIMENT builds are WRONG on purpose.
run() { TMX_NVCC_EXTRA="$1" python -m tweediemix_b200.build --force > /dev/null 2>&1; echo "== $1"; TMX_NVCC_EXTRA="$1" python tools/kbench.py --only attention --shapes Nq4096_Nk4096,Nq1024_Nk1024 2>&1 | grep "attention "; }
run "-DTMX_ATTN_POLY_EVERY=0"
run "-DTMX_ATTN_POLY_EVERY=2"
run "-DTMX_ATTN_POLY_EVERY=3"
run "-DTMX_ATTN_POLY_EVERY=0 -DTMX_ATTN_EXPERIMENT_NOEXP"
run "-DTMX_ATTN_POLY_EVERY=0 -DTMX_ATTN_EXPERIMENT_NOEXP -DTMX_ATTN_EXPERIMENT_NOMAX"
python -m tweediemix_b200.build --force > /dev/null 2>&1
