#!/bin/bash
P=tools/probe/tma_store_probe2
mkdir -p gpurun_out
{
for m in 0 1 2 3 4 5 6; do for w in 4 8 16; do
  if [ $m -ge 1 ] && [ $m -le 3 ] && [ $w -eq 16 ]; then d=1; else d=2; fi
  $P $m $w $d
done; done
} 2>&1 | tee gpurun_out/store_probe2_r2.txt
