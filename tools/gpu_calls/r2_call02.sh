#!/bin/bash
# Round 2, GPU call 2: the whole GPU suite on the adopted build, the new bench line, mesh run-length A/B, ncu evidence.
cd "$(dirname "$0")/../.."
rm -f gpurun_out/fullsize_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -150 > gpurun_out/r2b_gputests.log
tail -5 gpurun_out/r2b_gputests.log
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err
timeout 300 python bench.py --workload cfg2-hollow-sphere --no-per-config > gpurun_out/r2b_bench_cfg2.json 2>> gpurun_out/r2b_bench.err
echo "== mesh configs with the old run length (FTB_RUN_MAX=8) vs the new default"
FTB_RUN_MAX=8 bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" tree 2>&1 | tee gpurun_out/r2b_mesh_runmax8.log
bash tools/ab_bench.sh "cfg4-bunny cfg4-bunny-d12 cfg4-bunny-full-d14" tree 2>&1 | tee gpurun_out/r2b_mesh_default.log
echo "== e2e A/B on cfg2: host register / no staging"
FTB_HOST_REGISTER=1 bash tools/ab_bench.sh "cfg2-hollow-sphere cfg5-moon" tree 2>&1 | tee gpurun_out/r2b_e2e_register.log
FTB_NO_STAGING=1 bash tools/ab_bench.sh "cfg2-hollow-sphere cfg5-moon" tree 2>&1 | tee gpurun_out/r2b_e2e_nostaging.log
bash tools/ab_bench.sh "cfg2-hollow-sphere cfg5-moon" tree 2>&1 | tee gpurun_out/r2b_e2e_default.log
# ---- ncu: launch list of the default command, full sets of the top kernel on the headline and on cfg2
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config"
$CMD > gpurun_out/r2b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_launches.csv $CMD > gpurun_out/r2b_ncu1.log 2>&1
$CMD > gpurun_out/r2b_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -o gpurun_out/prof_r2b_repeat $CMD > gpurun_out/r2b_ncu2.log 2>&1
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config --workload cfg2-hollow-sphere"
$CMD2 > gpurun_out/r2b_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -o gpurun_out/prof_r2b_cfg2 $CMD2 > gpurun_out/r2b_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
