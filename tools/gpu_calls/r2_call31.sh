#!/bin/bash
# Round 2, GPU call 31: the common-origin bound table from 4 items on (tab4 = tree: the moon scene's four spheres, variants 0x230 / 0x250) against 8
# (notrig = the build of call 30); parity + golden + fuzz and the full-size moon / sample comparisons on the tree build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/ab_fast.py "cfg5-moon cfg1-sample" "notrig tab4" 5 2>&1 | tee gpurun_out/r2ae_table_min4_ab.txt
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2ae_parity.log
timeout 500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "moon or sample" 2>&1 | tail -5 | tee -a gpurun_out/r2ae_parity.log
cp gpurun_out/fullsize_parity.jsonl gpurun_out/r2ae_fullsize.jsonl 2>/dev/null
