#!/bin/bash
# Sustained (power-capped) lookup anatomy: whole kernel, gather only (1), resampling + stores only (2).
out=gpurun_out/power_lookup_anatomy.txt
: > $out
export RCB_USE_DEBUG_LIB=1
P="timeout 120 python tools/power_timeline.py --seconds 3 --what lookup"
for d in 0 1 2; do
  echo "### RCB_LOOKUP_DEBUG=$d" >> $out; RCB_LOOKUP_DEBUG=$d $P >> $out 2>&1
done
for d in 0 1 2; do
  echo "### burst RCB_LOOKUP_DEBUG=$d" >> $out; RCB_LOOKUP_DEBUG=$d timeout 100 python tools/time_lookup.py --reps 64 >> $out 2>&1
done
grep -E "###|t= 0\.0|t= 2\.[47]|lookup " $out
