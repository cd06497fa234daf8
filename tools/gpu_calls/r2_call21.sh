#!/bin/bash
# Round 2, GPU call 21: per-variant unrolling of the bound loops (1 for the large kernels) with and without the sign-built masks (t0u1 / tCv),
# the general bound loop on signs too (tCJv = tree), a house-family variant without the cube (v74a: repeat) - anchors: sinks2, tCu1 of call 20;
# parity file + fuzz on the tree build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python tools/ab_fast.py "cfg5-repeat cfg3-house cfg3-night-house cfg2-hollow-sphere cfg5-moon cfg4-bunny cfg4-bunny-full-d14 cfg1-sample" "sinks2 tCu1 t0u1 tCv tCJv tree v74a" 5 2>&1 | tee gpurun_out/r2u_unroll_signs_ab.txt
timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2u_parity.log
