#!/bin/bash
# Round 1's kernel (libraftcorr_b200_r1.so built from commit ed8c101), single-CTA MMAs against cta_group::2 pairs,
# in a burst and under sustained load: is halving the B bytes delivered per SM worth the pair's barrier traffic?
out=gpurun_out/r1_ncta_ab.txt
: > $out
export RCB_LIB_VARIANT=r1
for n in 1 2 1 2; do
  echo "### burst RCB_TC_NCTA=$n" >> $out
  RCB_TC_NCTA=$n timeout 120 python tools/time_build.py --mode bf16x3 --reps 10 2>&1 | grep -v Warn | grep build >> $out
done
for n in 1 2; do
  echo "### sustained build RCB_TC_NCTA=$n" >> $out
  RCB_TC_NCTA=$n timeout 120 python tools/power_timeline.py --seconds 3 --what build --mode bf16x3 2>&1 | grep -E "t= 2\.[47]" >> $out
  echo "### sustained step RCB_TC_NCTA=$n" >> $out
  RCB_TC_NCTA=$n timeout 120 python tools/power_timeline.py --seconds 3 --what step --mode bf16x3 2>&1 | grep -E "t= 2\.[47]" >> $out
done
cat $out
