#!/bin/bash
# Round 2, GPU call 22: ncu (full set, source) of the headline kernel and the house kernel of the tree build, to see what the hot set is made of now.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for w in cfg5-repeat cfg3-house; do
  FTB_AB_ARM=1 FTB_AB_TAG=tree timeout 200 python tools/ab_fast.py $w 2 > gpurun_out/r2v_plain_$w.log 2>&1 && \
  FTB_AB_ARM=1 FTB_AB_TAG=tree timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -o gpurun_out/prof_r2v_$w python tools/ab_fast.py $w 2 > gpurun_out/r2v_ncu_$w.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
