#!/bin/bash
# Runs bench.py for every (workload, library) pair on the SAME box and prints one line each:
#   workload lib ms/frame kernel_ms e2e_ms Mrays/s create_ms <scene create> bvh <device|host> <build kernels ms> <build total ms>
#   tools/ab_bench.sh "cfg2-hollow-sphere cfg3-night-house" "tree base exp1"
# "tree" is the in-tree library; any other name NAME selects ab/libftb_NAME.so (tools/ab_build.sh).
# Meant to be the command of one gpurun call; box-to-box variance is 1-2 %, so only same-box pairs compare.
workloads=${1:-cfg2-hollow-sphere}
libs=${2:-tree}
for w in $workloads; do
  for lib in $libs; do
    if [ "$lib" = tree ]; then unset FTB_LIB; else export FTB_LIB=$PWD/ab/libftb_$lib.so; fi
    timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-per-config --workload "$w" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
mi = d['e2e'].get('mesh_index') or {}
print('$w', '$lib', round(d['ms_per_step'], 4), round(d['roofline']['kernel_ms'], 4), round(d['e2e']['ms_per_step'], 4), round(d['value'], 1),
      'create_ms', round(d['e2e']['scene_create_ms'], 1), 'bvh', 'device' if mi.get('bvh_on_device') else 'host', round(mi.get('bvh_build_ms', 0), 2), round(mi.get('bvh_total_ms', 0), 1))"
  done
done
