#!/bin/bash
# Round 2, GPU call 23: leaf kind words carried by the item records (kw = tree) against the previous commit (tCJv), with the bound loops unrolled
# by 2 and 4 in every variant on top; parity file + golden + fuzz on the tree build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python tools/ab_fast.py "cfg5-repeat cfg3-house cfg3-night-house cfg2-hollow-sphere cfg5-moon cfg4-bunny cfg4-bunny-full-d14 cfg1-sample" "tCJv kw tree kwu2 kwu4" 5 2>&1 | tee gpurun_out/r2w_kindword_ab.txt
timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2w_parity.log
