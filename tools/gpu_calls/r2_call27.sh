#!/bin/bash
# Round 2, GPU call 27: packed FP32 (FFMA2 / FADD2) in the bound loops and in the world-to-model transform, each alone and together, against HEAD (base):
#   pkT = the common-origin table walked two items per step; pkTB = that + the general bound loop in pairs for every variant; pkTB1 = the general loop in
#   pkAllU = pkAll with the pair loops unrolled twice in every variant;
#   pairs only in the variants that built their mask from sign bits; pkAll / pkAll1 = pkTB / pkTB1 + origin and direction through a matrix row as one pair.
# Then parity + golden + fuzz on the tree build (= pkAll).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/ab_fast.py "cfg5-repeat cfg3-house cfg3-night-house cfg2-hollow-sphere cfg5-moon cfg4-bunny" "base pkT pkTB pkTB1 pkAll pkAll1 pkAllU" 3 2>&1 | tee gpurun_out/r2aa_packed_ab.txt
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2aa_parity.log
