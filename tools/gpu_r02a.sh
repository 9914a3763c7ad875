#!/bin/bash
# Round-2 pass A (1 GPU): new parity tests, kbench with same-box comparators, HV2_XCHG attention variant, bench line,
# ncu evidence for gn_fused_nhwc (application replay) and the k3 kernels.
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python -m tweediemix_b200.build > $OUT/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
grep -E "PARITY|adapter|passed|failed|rc=" $OUT/pytest_gpu.log | tail -30
timeout 600 python tools/kbench.py --compare > $OUT/kbench.txt 2>&1; echo "kbench rc=$?"
tail -60 $OUT/kbench.txt
# HV = 2 with the half-row max exchange (prebuilt variant library)
export TMX_LIB_PATH=$PWD/tweediemix_b200/lib/libtmx_hv2x.so
timeout 300 python -m pytest tests/test_gpu_attention.py -q -x > $OUT/pytest_attn_hv2x.log 2>&1; echo "attn hv2x default-variant tests rc=$?"
timeout 300 python tools/kbench.py --only attention --attn-variant 12 > $OUT/kbench_attn_hv2x.txt 2>&1; cat $OUT/kbench_attn_hv2x.txt | tail -6
timeout 300 python tools/attn_check.py 12 > $OUT/attn_check_hv2x.txt 2>&1; tail -3 $OUT/attn_check_hv2x.txt
unset TMX_LIB_PATH
timeout 900 python bench.py --steps 2 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
tail -c 3000 $OUT/bench.json; tail -3 $OUT/bench.err
# ncu: gn_fused_nhwc cannot be kernel-replayed (inter-CTA barrier) -> application replay with a reduced section set
NCUA="ncu --replay-mode application --clock-control none -f --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
timeout 900 $NCUA -k regex:gn_fused -s 4 -c 1 -o $OUT/gn_fused_c1280_hw1024 python tools/kbench.py --only groupnorm --shapes C1280_HW1024 --reps 1 > $OUT/ncu_gnf1.log 2>&1; echo "gn_fused c1280 rc=$?"
timeout 900 $NCUA -k regex:gn_fused -s 4 -c 1 -o $OUT/gn_fused_c640_hw4096 python tools/kbench.py --only groupnorm --shapes C640_HW4096 --reps 1 > $OUT/ncu_gnf2.log 2>&1; echo "gn_fused c640 rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:lora_delta -s 4 -c 1 -o $OUT/k3_lora_qkv_n1024 python tools/kbench.py --only routed --shapes lora_qkv_N1024 --reps 1 > $OUT/ncu_k3a.log 2>&1; echo "k3 lora rc=$?"
timeout 600 $NCU -k regex:routed_gemm -s 4 -c 1 -o $OUT/k3_gemm_kv_d1280 python tools/kbench.py --only routed --shapes kv_d1280 --reps 1 > $OUT/ncu_k3b.log 2>&1; echo "k3 gemm rc=$?"
ls -la $OUT
