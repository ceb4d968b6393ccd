#!/bin/bash
# A/B of lookup kernel variants (lanes per query, FFMA2): parity, burst and sustained timing.
# The variant libraries are built first, here: python -m raft_optical_flow_b200.build --variant lpq2b -DRCB_LOOKUP_LPQ=2
# (lpq2c: + -DRCB_LOOKUP_FFMA2=1, since removed from the source; lpq4c: -DRCB_LOOKUP_LPQ=4).
out=gpurun_out/lpq_ab2.txt
: > $out
echo "### parity (lpq2c library)" >> $out
RCB_LIB_VARIANT=lpq2c timeout 900 python -m pytest tests -m gpu -q -x -k "lookup or golden or oracle or edge or planned" 2>&1 | tail -3 >> $out
for v in "" lpq2b lpq2c lpq4c "" lpq2c; do
  echo "### variant='$v'" >> $out
  RCB_LIB_VARIANT=$v timeout 100 python tools/time_lookup.py --reps 128 2>&1 | grep lookup >> $out
  RCB_LIB_VARIANT=$v timeout 120 python tools/power_timeline.py --seconds 3 --what step 2>&1 | grep -E "t= 2\.[47]" >> $out
done
cat $out
