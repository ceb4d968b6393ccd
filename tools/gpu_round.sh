#!/bin/bash
# One GPU session: build-kernel timings per mode (release + debug anatomy), then (part of) the GPU test suite.
mkdir -p gpurun_out
{
for mode in f16f8 bf16x3 bf16; do python tools/time_build.py --mode $mode --reps 10 2>&1 | grep -v Warn; done
python tools/time_build.py --mode f16f8 --reps 10 --between lookups 2>&1 | grep -v Warn
export RCB_USE_DEBUG_LIB=1
for mode in ${MODES:-f16f8}; do for skip in ${SKIPS:-0 32 7 39}; do
  RCB_TC_PROF=1 RCB_TC_DEBUG_SKIP=$skip python tools/time_build.py --mode $mode --reps 10 2>&1 | grep -v Warn
done; done
unset RCB_USE_DEBUG_LIB
} 2>&1 | tee gpurun_out/build_times.txt
timeout 1500 python -m pytest tests -m gpu -q -x -k "${TESTK:-not zzz}" 2>&1 | tail -15 | tee gpurun_out/gputest.txt
