#!/bin/bash
# Round 2, GPU call 33: the final build of the round (packed FP32 bound tests, packed two-child box tests + 64-byte node records, angle-free Oren-Nayar) - whole GPU suite (full-size parity records), the bench lines (default = headline + per_config, reference arm),
# the launch list and the full ncu set of the headline kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/fullsize_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -40 > gpurun_out/r2ag_gputests.log
tail -4 gpurun_out/r2ag_gputests.log
timeout 900 python bench.py > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err; tail -2 gpurun_out/r2ag_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2ag_bench_reference.json 2>> gpurun_out/r2ag_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config"
$CMD > gpurun_out/r2ag_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2ag_launches.csv $CMD > gpurun_out/r2ag_ncu1.log 2>&1
$CMD > gpurun_out/r2ag_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -o gpurun_out/prof_r2ag_repeat $CMD > gpurun_out/r2ag_ncu2.log 2>&1
ls -la gpurun_out/prof_r2ag_repeat.ncu-rep
head -c 600 gpurun_out/r2ag_bench.json
