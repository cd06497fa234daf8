#!/bin/bash
# Round 2, GPU call 7: mesh walk chosen by mesh size (per-lane below 32 k triangles, packet above), device-built index vs host; create phases; GPU suite.
cd "$(dirname "$0")/../.."
echo "== device-built index (default)"
bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" "tree" 2>&1 | tee gpurun_out/r2g_mesh_device.log
echo "== host-built index (FTB_HOST_BVH=1)"
FTB_HOST_BVH=1 bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" "tree" 2>&1 | tee gpurun_out/r2g_mesh_host.log
echo "== scene create phases"
FTB_VERBOSE=1 timeout 120 python tools/scene_create_time.py 2>&1 | tee gpurun_out/r2g_create_device.log
FTB_HOST_BVH=1 FTB_VERBOSE=1 timeout 120 python tools/scene_create_time.py cfg4-bunny-full-d14 2>&1 | tee gpurun_out/r2g_create_host.log
rm -f gpurun_out/fullsize_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -60 > gpurun_out/r2g_gputests.log
tail -8 gpurun_out/r2g_gputests.log
