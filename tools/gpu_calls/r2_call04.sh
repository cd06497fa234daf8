#!/bin/bash
# Round 2, GPU call 4: fused solidCylinder leaf against the previous build (ab/libftb_bvh2.so = the commit before it), same box; GPU suite.
cd "$(dirname "$0")/../.."
bash tools/ab_bench.sh "cfg3-house cfg3-night-house cfg5-repeat cfg2-hollow-sphere" "bvh2 tree" 2>&1 | tee gpurun_out/r2d_solidcyl_ab.log
rm -f gpurun_out/fullsize_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -60 > gpurun_out/r2d_gputests.log
tail -5 gpurun_out/r2d_gputests.log
