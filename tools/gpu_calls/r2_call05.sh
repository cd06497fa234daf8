#!/bin/bash
# Round 2, GPU call 5: warp-packet mesh traversal against the per-lane binary BVH walk (ab/libftb_bvh2.so), same box; GPU suite.
cd "$(dirname "$0")/../.."
for rm in 1 8; do
  echo "== FTB_RUN_MAX=$rm"
  FTB_RUN_MAX=$rm bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" "bvh2 tree pk5 pk4"
done 2>&1 | tee gpurun_out/r2e_packet_ab.log
rm -f gpurun_out/fullsize_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -60 > gpurun_out/r2e_gputests.log
tail -5 gpurun_out/r2e_gputests.log
