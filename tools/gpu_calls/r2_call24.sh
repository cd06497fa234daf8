#!/bin/bash
# Round 2, GPU call 24: block masks for primary rays - one loop over the set bits for every tabled ray (bmA = tree), or that loop for primaries and the
# descending loop for shadow rays (bmB) - against the previous commit (kw); parity file + golden + fuzz on the tree build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python tools/ab_fast.py "cfg5-repeat cfg3-house cfg3-night-house cfg2-hollow-sphere cfg5-moon cfg4-bunny" "kw bmA bmB" 5 2>&1 | tee gpurun_out/r2x_blockmask_ab.txt
timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2x_parity.log
