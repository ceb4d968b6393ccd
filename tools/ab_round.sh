#!/bin/bash
# A/B timing of build-kernel variants inside ONE GPU session (libraftcorr_b200_<name>.so from build.py --variant).
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in ${VARIANTS:-a1 a3}; do
  for mode in ${MODES:-f16f8}; do
  for skip in ${SKIPS:-0}; do
    echo -n "variant $v nstage=${RCB_TC_NSTAGE:-max}: "
    RCB_LIB_VARIANT=$v RCB_TC_DEBUG_SKIP=$skip python tools/time_build.py --mode $mode --reps 10 2>&1 | grep "^cfg"
  done; done
done
done
} 2>&1 | tee -a gpurun_out/ab_round.txt
for v in ${TESTV:-}; do
RCB_LIB_VARIANT=$v timeout 900 python -m pytest tests -m gpu -q -x -k "${TESTK:-corrblock_vs_oracle or full_size or guard}" 2>&1 | tail -3 | tee -a gpurun_out/ab_round.txt
done
