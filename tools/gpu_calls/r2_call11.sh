#!/bin/bash
# Round 2, GPU call 11: one-copy leaf intersectors (ab/libftb_onecopy.so) vs the tree on the house family; mesh configs on the final
# policy (device-built index for large meshes); GPU suite; ncu of the packet walk on the 355 k-triangle mesh.
cd "$(dirname "$0")/../.."
bash tools/ab_bench.sh "cfg3-house cfg3-night-house cfg5-repeat" "tree onecopy" 2>&1 | tee gpurun_out/r2k_onecopy_ab.log
bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" "tree" 2>&1 | tee gpurun_out/r2k_mesh_final.log
FTB_HOST_BVH=1 bash tools/ab_bench.sh "cfg4-bunny-full-d14" "tree" 2>&1 | tee -a gpurun_out/r2k_mesh_final.log
FTB_VERBOSE=1 timeout 120 python tools/scene_create_time.py cfg4-bunny-full-d14 2>&1 | grep -v "bvh build" | tee gpurun_out/r2k_create_device.log
rm -f gpurun_out/fullsize_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -60 > gpurun_out/r2k_gputests.log
tail -4 gpurun_out/r2k_gputests.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config --workload cfg4-bunny-full-d14"
$CMD > gpurun_out/r2k_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -o gpurun_out/prof_r2k_mesh $CMD > gpurun_out/r2k_ncu.log 2>&1
ls -la gpurun_out/prof_r2k_mesh.ncu-rep
