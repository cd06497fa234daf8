#!/bin/bash
# Round 2, GPU call 29: mesh walks - both child boxes of a node as packed pairs (box2), + 64-byte node records with the links in the fourth row (node),
# + the packet walk asking for both children's records while it tests the current node (nodepf = tree), against the packed build of call 28 (pk2);
# parity + golden + fuzz on the tree build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/ab_fast.py "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" "pk2 box2 node nodepf" 5 2>&1 | tee gpurun_out/r2ac_mesh_nodes_ab.txt
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2ac_parity.log
