#!/bin/bash
# Anatomy of the tcgen05 build kernel: role cycle counters, epilogue phase timers and phase-skip timings
# (debug library only: python -m raft_optical_flow_b200.build --debug).
export RCB_USE_DEBUG_LIB=1
mkdir -p gpurun_out
{
for mode in ${MODES:-f16f8}; do
for skip in ${SKIPS:-0 32 7 39}; do
  RCB_TC_PROF=1 RCB_TC_DEBUG_SKIP=$skip python tools/time_build.py --mode $mode --reps 10 2>&1 | grep -v Warning
done
done
} 2>&1 | tee gpurun_out/build_anatomy_r2.txt
