#!/bin/bash
# Round 2, GPU call 30: Oren-Nayar without forming the angles in the FP32 build (notrig = tree) against the literal angle form (trig) on the moon frame;
# parity + golden + the full-size moon comparison on the tree build; then an ncu set (full, source) of the moon kernel, which has never been profiled.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/ab_fast.py "cfg5-moon" "trig notrig" 5 2>&1 | tee gpurun_out/r2ad_rough_trig_ab.txt
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2ad_parity.log
timeout 500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k moon 2>&1 | tail -5 | tee -a gpurun_out/r2ad_parity.log
cp gpurun_out/fullsize_parity.jsonl gpurun_out/r2ad_fullsize_moon.jsonl 2>/dev/null
FTB_AB_ARM=1 FTB_AB_TAG=tree timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -o gpurun_out/prof_r2ad_moon python tools/ab_fast.py cfg5-moon 2 > gpurun_out/r2ad_ncu_moon.log 2>&1
ls -la gpurun_out/*r2ad*
