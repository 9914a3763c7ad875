#!/bin/bash
# Polynomial-exp fraction sweep for the attention kernel (GPU box).
run() { TMX_NVCC_EXTRA="$1" python -m tweediemix_b200.build --force > /dev/null 2>&1; echo "== $1"; TMX_NVCC_EXTRA="$1" python tools/kbench.py --only attention --shapes Nq4096_Nk4096,Nq1024_Nk1024 2>&1 | grep "attention "; }
for pe in 0 2 3 5 8; do run "-DTMX_ATTN_POLY_EVERY=$pe"; done
python -m tweediemix_b200.build --force > /dev/null 2>&1
