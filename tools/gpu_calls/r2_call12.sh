#!/bin/bash
# Round 2, GPU call 12: the tree against the build of call 4 (ab/libftb_c4.so: before the packet walk / device build went in) on the
# non-mesh variants: nothing but the mesh variants may have changed.
cd "$(dirname "$0")/../.."
bash tools/ab_bench.sh "cfg3-house cfg3-night-house cfg5-repeat cfg2-hollow-sphere cfg5-moon cfg1-sample" "c4 tree" 2>&1 | tee gpurun_out/r2l_regression_ab.log
