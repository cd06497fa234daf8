#!/bin/bash
# Round 2, GPU call 13: 4 CTAs/SM (128 registers, no spills) and 256-sample blend units against the tree.
cd "$(dirname "$0")/../.."
bash tools/ab_bench.sh "cfg5-repeat cfg5-moon cfg3-house cfg3-night-house cfg2-hollow-sphere" "tree mb4 cap256" 2>&1 | tee gpurun_out/r2m_mb4_cap256_ab.log
