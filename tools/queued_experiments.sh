#!/bin/bash
# The experiments DESIGN.md §8 lists as "built but not yet measured", as two steps:
#   tools/queued_experiments.sh build        (here, no GPU: builds ab/libftb_{cursor,net,cubebf,tight,all4}.so, prints SASS sizes / spills)
#   gpurun --timeout 1500 -- 'bash tools/queued_experiments.sh run 2>&1 | tee gpurun_out/queued.log'
# "run" first checks every library against the oracle (the FP32 / FP64 parity tests and the fuzz scenes go through
# FTB_LIB), then prints same-box timings next to the in-tree build.  A switch is adopted only if parity is green and it
# wins on the workloads it targets without losing > 1 % elsewhere (box-to-box variance is 1-2 %: compare within one run).
set -e
cd "$(dirname "$0")/.."
LIBS="cursor net cubebf tight all4"
case "$1" in
build)
  bash tools/ab_build.sh cursor "-DFTB_CURSOR_SMEM=1"
  bash tools/ab_build.sh net    "-DFTB_PAIR_NETWORK=1"
  bash tools/ab_build.sh cubebf "-DFTB_CUBE_BRANCHFREE=1"
  bash tools/ab_build.sh tight  "-DFTB_TABLE_TIGHT_SLACK=1"
  bash tools/ab_build.sh all4   "-DFTB_CURSOR_SMEM=1 -DFTB_PAIR_NETWORK=1 -DFTB_CUBE_BRANCHFREE=1 -DFTB_TABLE_TIGHT_SLACK=1"
  # mesh frames are bound by the latency of dependent node fetches from L2 (4 % of the rays walk ~90 nodes, tools/bvh_quality.cpp):
  # more resident warps at the price of spills -- for the mesh workloads only (cfg2 / cfg3 measured slower with 6 and 8)
  bash tools/ab_build.sh mb6    "-DFTB_MIN_BLOCKS=6"
  bash tools/ab_build.sh mb8    "-DFTB_MIN_BLOCKS=8"
  bash tools/ab_build.sh mb6cur "-DFTB_MIN_BLOCKS=6 -DFTB_CURSOR_SMEM=1"
  # lanes idle when both blend-ring slots still wait for a long sample: with the traversal's cost variance a third / fourth
  # unit in flight may pay on meshes although it does not elsewhere
  bash tools/ab_build.sh rs3    "-DFTB_RING_SLOTS=3"
  bash tools/ab_build.sh rs4    "-DFTB_RING_SLOTS=4"
  # every general CSG item of the bundled scenes is `subtract (solidCylinder) (sphere)`: with operands that may be runs of
  # leaves it becomes a register-resident pair and the house family no longer needs the general evaluator (variant 0x34b)
  AB_MAKE_ARGS="EXTRA=-DFTB_PAIR_GROUPS=1 'F32_FEATS=0x3ff 0x000 0x100 0x209 0x34b 0x3c3 0x050 0x030 0x004'" bash tools/ab_build.sh groups ""
  AB_MAKE_ARGS="EXTRA=-DFTB_PAIR_GROUPS=1 'F32_FEATS=0x3ff 0x000 0x100 0x209 0x34b 0x3c3 0x050 0x030 0x004'" bash tools/ab_build.sh groupscur "-DFTB_CURSOR_SMEM=1"
  ;;
run)
  for lib in $LIBS; do
    echo "== parity with ab/libftb_$lib.so"
    FTB_LIB=$PWD/ab/libftb_$lib.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -n 2
  done
  bash tools/ab_bench.sh "cfg2-hollow-sphere cfg3-house cfg3-night-house cfg4-bunny-full-d14 cfg5-moon cfg5-repeat" "tree $LIBS"
  bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" "tree mb6 mb8 mb6cur rs3 rs4"
  # run length of the sample dealing (host side, no rebuild): 16 spp deals runs of 2 samples; single samples keep a warp on fewer pixels
  for rm in 1 2 8; do echo "== FTB_RUN_MAX=$rm"; FTB_RUN_MAX=$rm bash tools/ab_bench.sh "cfg4-bunny-full-d14 cfg3-house cfg5-moon" "tree"; done
  for lib in groups groupscur; do
    echo "== parity with ab/libftb_$lib.so"
    FTB_LIB=$PWD/ab/libftb_$lib.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -n 2
  done
  bash tools/ab_bench.sh "cfg3-house cfg3-night-house cfg5-repeat cfg2-hollow-sphere" "tree groups groupscur"
  ;;
*) echo "usage: $0 build|run"; exit 2;;
esac
