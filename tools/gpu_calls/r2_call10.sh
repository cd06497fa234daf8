#!/bin/bash
# Round 2, GPU call 10 (8 GPUs): the sweep config at 8 and 4 ranks.
cd "$(dirname "$0")/../.."
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2j_bench_n$n.json 2> gpurun_out/r2j_bench_n$n.err; echo "n=$n rc=$?"; tail -3 gpurun_out/r2j_bench_n$n.err | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 3 --workload cfg2-hollow-sphere --no-inprocess > gpurun_out/r2j_bench_n8_cfg2.json 2> gpurun_out/r2j_bench_n8_cfg2.err; echo "cfg2 rc=$?"
