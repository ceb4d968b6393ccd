#!/bin/bash
# Stagger sweep: CTA groups start their patch sweep at different patches (alias = -S).
P=tools/probe/tma_store_probe
mkdir -p gpurun_out
{
for w in 4 8; do for S in 1 2 4 7 8 14 28 56; do $P 128 $w 2 -$S; done; done
for S in 1 8 56; do $P 256 4 2 -$S; done
} 2>&1 | tee gpurun_out/store_probe_stagger_r2.txt
