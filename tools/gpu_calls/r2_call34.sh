#!/bin/bash
# Round 2, GPU call 34: ncu sets (full, source) of the final build's kernels on the 355 k-triangle mesh (packet walk on 64-byte node records) and on the
# hollow-sphere frame, the two kernels VERDICT r1 singled out.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for w in cfg4-bunny-full-d14 cfg2-hollow-sphere; do
  FTB_AB_ARM=1 FTB_AB_TAG=tree timeout 200 python tools/ab_fast.py $w 2 > gpurun_out/r2ah_plain_$w.log 2>&1 && \
  FTB_AB_ARM=1 FTB_AB_TAG=tree timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -o gpurun_out/prof_r2ah_$w python tools/ab_fast.py $w 2 > gpurun_out/r2ah_ncu_$w.log 2>&1
done
ls -la gpurun_out/prof_r2ah_*.ncu-rep
