#!/bin/bash
# Sustained-load clock/power attribution for the build kernel: runs tools/time_build.py in a long loop under several
# debug-skip settings while nvidia-smi samples clocks and power; prints the median of the samples taken under load.
#   bash tools/power_probe.sh
run() {
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 50 > /tmp/clk_$1.csv & NSPID=$!
  eval "$2 python tools/time_build.py --reps 5000 $3" | tail -1
  kill $NSPID; wait $NSPID 2>/dev/null
  python - "$1" <<'PY'
import sys, statistics
rows = [l.split(",") for l in open(f"/tmp/clk_{sys.argv[1]}.csv") if l.strip()]
rows = [(float(a), float(b)) for a, b in rows]
load = [r for r in rows if r[1] > 0.8 * max(x[1] for x in rows)]
print(f"   {sys.argv[1]:28s} median under load: {statistics.median(r[0] for r in load):6.0f} MHz {statistics.median(r[1] for r in load):6.0f} W ({len(load)} samples)")
PY
}
run bf16x3_ncta2 "" ""
run bf16x3_ncta1 "RCB_TC_NCTA=1" ""
run bf16_ncta2 "" "--mode bf16"
run bf16x3_no_mma "RCB_TC_DEBUG_SKIP=32" ""
run bf16x3_no_l0_stores "RCB_TC_DEBUG_SKIP=1" ""
run bf16x3_no_stores "RCB_TC_DEBUG_SKIP=7" ""
run bf16x3_no_loads_no_mma "RCB_TC_DEBUG_SKIP=48" ""
