#!/bin/bash
# Attention tuning sweeps (GPU box).  -DTMX_ATTN_POLY_EVERY=n: one pair of exponentials in n on the FMA pipe (0 = all MUFU);
# -DTMX_ATTN_EXPERIMENT_NOEXP / _NOMAX: bottleneck-location builds whose RESULTS ARE WRONG on purpose.
run() { TMX_NVCC_EXTRA="$1" python -m tweediemix_b200.build --force > /dev/null 2>&1; echo "== $1"; TMX_NVCC_EXTRA="$1" python tools/kbench.py --only attention --shapes Nq4096_Nk4096,Nq1024_Nk1024 2>&1 | grep "attention "; }
for pe in 0 2 3 4 6; do run "-DTMX_ATTN_POLY_EVERY=$pe"; done
run "-DTMX_ATTN_POLY_EVERY=0 -DTMX_ATTN_EXPERIMENT_NOEXP"
python -m tweediemix_b200.build --force > /dev/null 2>&1
