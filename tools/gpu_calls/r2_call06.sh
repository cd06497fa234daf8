#!/bin/bash
# Round 2, GPU call 6: packet traversal with the cull on pop; mesh index built on the device (PLOC) against the host SAH build; GPU suite.
cd "$(dirname "$0")/../.."
export FTB_VERBOSE=1
echo "== host-built index (FTB_HOST_BVH=1): per-lane walk (bvh2) vs packet walk (tree)"
FTB_HOST_BVH=1 bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" "bvh2 tree" 2>&1 | tee gpurun_out/r2f_packet_hostbvh.log
echo "== device-built index (default)"
bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" "tree" 2>&1 | tee gpurun_out/r2f_packet_devicebvh.log
FTB_RUN_MAX=8 bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12" "tree" 2>&1 | tee -a gpurun_out/r2f_packet_devicebvh.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-per-config --workload cfg4-bunny-full-d14 2>&1 | tail -3 | cut -c1-600
rm -f gpurun_out/fullsize_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -60 > gpurun_out/r2f_gputests.log
tail -8 gpurun_out/r2f_gputests.log
