#!/bin/bash
mkdir -p gpurun_out
export RCB_USE_DEBUG_LIB=1
{
for mode in ${MODES:-f16f8}; do for skip in ${SKIPS:-16 48 20 4}; do
  RCB_TC_PROF=1 RCB_TC_DEBUG_SKIP=$skip python tools/time_build.py --mode $mode --reps 10 2>&1 | grep -v Warn
done; done
} 2>&1 | tee gpurun_out/build_times2.txt
unset RCB_USE_DEBUG_LIB
timeout 900 python -m pytest tests -m gpu -q -x -k "${TESTK:-from_packed or alternate_block_backward}" 2>&1 | tail -8 | tee gpurun_out/gputest2.txt
