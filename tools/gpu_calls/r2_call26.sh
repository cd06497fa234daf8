#!/bin/bash
# Round 2, GPU call 26 (8 GPUs): every config at 1 / 2 / 4 / 8 GPUs through ftb_render(n_gpus = N) in one process (tools/scale_inprocess.py),
# then the headline bench line at 8 ranks under torchrun (value, e2e, frame checks, e2e_inprocess).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/scale_inprocess.py --steps 3 > gpurun_out/r2z_scale_inprocess.jsonl 2> gpurun_out/r2z_scale_inprocess.md; echo "sweep rc=$?"; tail -12 gpurun_out/r2z_scale_inprocess.md
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2z_bench_n8.json 2> gpurun_out/r2z_bench_n8.err; echo "n8 rc=$?"; tail -2 gpurun_out/r2z_bench_n8.err | cut -c1-300
head -c 1500 gpurun_out/r2z_bench_n8.json
