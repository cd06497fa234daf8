#!/bin/bash
# Round 2, GPU call 19: the segmented fold sized by the unit (fold2), push-front pair sinks (sinks), the packed ray sink on top (sinks2 = tree),
# the same at 5 CTAs/SM (s2mb5), unroll 2 (u2), pixel masks without the shadow table (masks2) - all against the previous commit (base);
# then ncu of base and masks on cfg2 (why did the masks lose?), and the parity file on the tree build.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python tools/ab_fast.py "cfg5-repeat cfg3-house cfg3-night-house cfg2-hollow-sphere cfg5-moon cfg4-bunny cfg4-bunny-full-d14" "base fold2 sinks sinks2 tree s2mb5 u2 masks2" 5 2>&1 | tee gpurun_out/r2s_sinks_ab.txt
for lib in base masks sinks2; do
  FTB_AB_ARM=1 FTB_AB_TAG=$lib FTB_LIB=$PWD/ab/libftb_$lib.so timeout 300 ncu --set full --clock-control none -k regex:render_kernel -s 2 -c 1 -o gpurun_out/prof_r2s_cfg2_$lib python tools/ab_fast.py cfg2-hollow-sphere 2 > gpurun_out/r2s_ncu_$lib.log 2>&1
done
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2s_parity.log
ls -la gpurun_out/*.ncu-rep
