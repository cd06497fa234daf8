#!/bin/bash
# Round 2, GPU call 16: the wavefront arm against the megakernel (frames must be bit-identical), the parity suite on the wavefront arm,
# and the "A has no crossings" shortcut of the CSG pairs against the build of call 14 (ab/libftb_c14.so).
cd "$(dirname "$0")/../.."
timeout 900 python tools/wavefront_ab.py 2>&1 | tee gpurun_out/r2p_wavefront_ab.log
FTB_WAVEFRONT=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2p_wavefront_tests.log
bash tools/ab_bench.sh "cfg2-hollow-sphere cfg3-house cfg3-night-house cfg5-repeat" "c14 tree" 2>&1 | tee gpurun_out/r2p_amiss_ab.log
