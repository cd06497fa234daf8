#!/bin/bash
# Round 2, GPU call 17 (2 GPUs): rotated tile dealing - multi-process bench with frame checks, in-process n_gpus test; on GPU 0: how far the
# wavefront frames are from the megakernel's, and ncu of one wavefront frame of cfg3-house next to the megakernel's launch.
cd "$(dirname "$0")/../.."
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err; echo "rc=$?"; tail -2 gpurun_out/r2q_bench_n2.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg2-hollow-sphere > gpurun_out/r2q_bench_n2_cfg2.json 2>> gpurun_out/r2q_bench_n2.err; echo "rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "in_process_multi_gpu or sharding or banded" 2>&1 | tail -3
export CUDA_VISIBLE_DEVICES=0
timeout 600 python tools/wavefront_ab.py cfg3-house cfg2-hollow-sphere 2>&1 | tee gpurun_out/r2q_wavefront_diff.log
CMD="python tools/wavefront_ab.py cfg3-house"
FTB_AB_ARM=1 FTB_WAVEFRONT=1 $CMD > gpurun_out/r2q_plain_wf.log 2>&1 && FTB_AB_ARM=1 FTB_WAVEFRONT=1 ncu --set full --clock-control none -k regex:wf_ -s 30 -c 6 -o gpurun_out/prof_r2q_wavefront_house $CMD > gpurun_out/r2q_ncu_wf.log 2>&1
FTB_AB_ARM=1 $CMD > gpurun_out/r2q_plain_mk.log 2>&1 && FTB_AB_ARM=1 ncu --set full --clock-control none -k regex:render_kernel -s 5 -c 1 -o gpurun_out/prof_r2q_megakernel_house $CMD > gpurun_out/r2q_ncu_mk.log 2>&1
ls -la gpurun_out/prof_r2q_*.ncu-rep
