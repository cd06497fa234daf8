#!/bin/bash
# Round 2, GPU call 18: per-pixel primary candidate masks and the segmented fold against the previous commit (ab/libftb_base.so), each alone
# and together (tree); FTB_RUN_MAX=4; 96-thread CTAs x 6; then the parity file and the golden fixtures on the tree build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/ab_fast.py "cfg5-repeat cfg3-house cfg3-night-house cfg2-hollow-sphere cfg5-moon cfg1-sample cfg4-bunny" "base fold masks tree tree@FTB_RUN_MAX=4 b96" 5 2>&1 | tee gpurun_out/r2r_masks_fold_ab.txt
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2r_parity.log
