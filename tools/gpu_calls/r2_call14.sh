#!/bin/bash
# Round 2, GPU call 14: per-variant CTA count (4 for the house family) and blend-unit size (256 for simple scenes) against the
# build of call 4; GPU suite; the final bench lines; ncu of the headline kernel.
cd "$(dirname "$0")/../.."
bash tools/ab_bench.sh "cfg5-repeat cfg5-moon cfg3-house cfg3-night-house cfg2-hollow-sphere cfg1-sample" "c4 tree" 2>&1 | tee gpurun_out/r2n_final_ab.log
rm -f gpurun_out/fullsize_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -60 > gpurun_out/r2n_gputests.log
tail -4 gpurun_out/r2n_gputests.log
timeout 900 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -2 gpurun_out/r2n_bench.err
timeout 300 python bench.py --workload cfg2-hollow-sphere --no-per-config > gpurun_out/r2n_bench_cfg2.json 2>> gpurun_out/r2n_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2n_bench_reference.json 2>> gpurun_out/r2n_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config"
$CMD > gpurun_out/r2n_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2n_launches.csv $CMD > gpurun_out/r2n_ncu1.log 2>&1
$CMD > gpurun_out/r2n_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -o gpurun_out/prof_r2n_repeat $CMD > gpurun_out/r2n_ncu2.log 2>&1
ls -la gpurun_out/prof_r2n_repeat.ncu-rep
