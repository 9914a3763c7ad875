set -x
timeout 300 python -m pytest tests/test_gpu_attention.py -m gpu -q -x 2>&1 | tail -3
python tools/kbench.py --only attention 2>&1 | grep attention
for pe in 0 4 8; do
  TMX_NVCC_EXTRA="-DTMX_ATTN_POLY_EVERY=$pe" python -m tweediemix_b200.build --force > /dev/null 2>&1
  echo "POLY_EVERY=$pe"; TMX_NVCC_EXTRA="-DTMX_ATTN_POLY_EVERY=$pe" python tools/kbench.py --only attention 2>&1 | grep attention
done
