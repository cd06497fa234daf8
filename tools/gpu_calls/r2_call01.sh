#!/bin/bash
# Round 2, GPU call 1: timings only for every queued switch (parity is run afterwards for the ones that win).
cd "$(dirname "$0")/../.."
bash tools/ab_bench.sh "cfg2-hollow-sphere cfg3-house cfg3-night-house cfg5-moon cfg5-repeat" "tree cursor net cubebf tight all4"
bash tools/ab_bench.sh "cfg3-house cfg3-night-house cfg5-repeat cfg2-hollow-sphere" "groups groupscur"
bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" "tree cursor mb6 mb8 mb6cur rs3 rs4"
for rm in 1 2; do echo "== FTB_RUN_MAX=$rm"; FTB_RUN_MAX=$rm bash tools/ab_bench.sh "cfg4-bunny-full-d14 cfg3-house cfg5-moon" "tree"; done
