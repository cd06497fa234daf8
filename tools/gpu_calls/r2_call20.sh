#!/bin/bash
# Round 2, GPU call 20: on top of the sinks (ab/libftb_sinks2.so = previous commit): quad fold + planeT (t0), the CSG rule walk as one rotating loop (tB),
# the bound table's mask by sign + funnel shift (tC), both (tBC = tree), with the bound loops unrolled by 2 / 1; parity file on the tree build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python tools/ab_fast.py "cfg5-repeat cfg3-house cfg3-night-house cfg2-hollow-sphere cfg5-moon cfg4-bunny" "sinks2 t0 tB tC tBC tBCu2 tCu2 tCu1" 5 2>&1 | tee gpurun_out/r2t_walk_signs_ab.txt
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2t_parity.log
