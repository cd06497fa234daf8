#!/bin/bash
# Round 2, GPU call 9 (2 GPUs): the multi-process bench path (banded e2e, e2e_inprocess, frame checks) and the in-process n_gpus test.
cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "rc=$?"; tail -c 3000 gpurun_out/r2i_bench_n2.json; tail -5 gpurun_out/r2i_bench_n2.err
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg2-hollow-sphere > gpurun_out/r2i_bench_n2_cfg2.json 2>> gpurun_out/r2i_bench_n2.err; echo "rc=$?"; tail -c 2500 gpurun_out/r2i_bench_n2_cfg2.json
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "in_process_multi_gpu or banded or pageable" 2>&1 | tail -5
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 0 | cut -c1-300
