#!/bin/bash
# Round 2, GPU call 3: 4-wide BVH with four-slot leaf blocks against the binary BVH of the previous commit (same box), at 4 / 5 / 6 CTAs
# per SM and with runs of 1 / 8 samples; then the whole GPU suite on the new build.
cd "$(dirname "$0")/../.."
for rm in 1 8; do
  echo "== FTB_RUN_MAX=$rm"
  FTB_RUN_MAX=$rm bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" "bvh2 tree b4mb5 b4mb4"
done 2>&1 | tee gpurun_out/r2c_mesh_ab.log
echo "== e2e with the adaptive band count"
bash tools/ab_bench.sh "cfg2-hollow-sphere cfg3-house cfg5-moon" tree 2>&1 | tee gpurun_out/r2c_e2e.log
rm -f gpurun_out/fullsize_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -60 > gpurun_out/r2c_gputests.log
tail -5 gpurun_out/r2c_gputests.log
