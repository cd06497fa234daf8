#!/bin/bash
# Round 2, GPU call 32 (2 GPUs): the bench line at 2 ranks under torchrun on the final build (frame checks against the 1-GPU render, e2e, e2e_inprocess),
# and the multi-GPU tests that a 1-GPU box skips.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/r2af_bench_n2.json 2> gpurun_out/r2af_bench_n2.err; echo "n2 rc=$?"; tail -2 gpurun_out/r2af_bench_n2.err | cut -c1-300
tail -c 1500 gpurun_out/r2af_bench_n2.json
timeout 200 python -m pytest tests -m gpu -q -x -k "multi_gpu or in_process or shard" 2>&1 | tail -4 | tee gpurun_out/r2af_multigpu_tests.log
