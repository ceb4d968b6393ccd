#!/bin/bash
# Sustained (power-capped) anatomy of the build: which part draws the power?  RCB_DEBUG library, skip bits of
# corr_build_tc.cu: 16 no B loads, 32 no MMAs, 64 no tmem_ld, 8 no level-0 stores, 128 no epilogue at all.
out=gpurun_out/power_build_anatomy.txt
: > $out
export RCB_USE_DEBUG_LIB=1
for skip in 0 32 48 8 128 176; do
  echo "### RCB_TC_DEBUG_SKIP=$skip" >> $out
  RCB_TC_DEBUG_SKIP=$skip timeout 120 python tools/power_timeline.py --seconds 3 --what build 2>&1 | grep -E "t= 0\.0|t= 2\.[47]" >> $out
done
cat $out
