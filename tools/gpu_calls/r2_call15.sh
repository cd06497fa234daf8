#!/bin/bash
# Round 2, GPU call 15 (8 GPUs): the final build on the sweep config at 2 / 4 / 8 ranks, cfg2 and cfg5-moon at 8.
cd "$(dirname "$0")/../.."
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2o_bench_n$n.json 2> gpurun_out/r2o_bench_n$n.err; echo "n=$n rc=$?"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 8 --steps 10 --warmup 3 --workload cfg5-moon > gpurun_out/r2o_bench_n8_moon.json 2> gpurun_out/r2o_bench_n8_moon.err; echo "moon rc=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "in_process_multi_gpu" 2>&1 | tail -2
