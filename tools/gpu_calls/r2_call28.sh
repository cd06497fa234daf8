#!/bin/bash
# Round 2, GPU call 28: the matrix of an item's (first) leaf loaded together with the item record (rows) against the packed build without it (pk2),
# both against HEAD (base); parity + golden + fuzz on the tree build (= rows).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/ab_fast.py "cfg5-repeat cfg3-house cfg3-night-house cfg2-hollow-sphere cfg5-moon cfg4-bunny cfg1-sample" "base pk2 rows" 3 2>&1 | tee gpurun_out/r2ab_item_rows_ab.txt
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2ab_parity.log
