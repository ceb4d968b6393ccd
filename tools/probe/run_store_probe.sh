#!/bin/bash
# Sweep of tools/probe/tma_store_probe on one B200: which store pattern of the build epilogue sustains what.
P=tools/probe/tma_store_probe
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "# kernel-like pattern (two interleaved 256-byte runs, lockstep)"
for w in 4 8; do for d in 2 4; do $P 128 $w $d -1; done; done
echo "# runs pattern: run_bytes x streams"
for w in 4 8; do
for rs in "256 2" "512 2" "1024 2" "2048 2" "256 1" "512 1" "2048 1"; do $P 128 $w 2 0 28672 $rs; done
done
echo "# full walk of a plane per warp (id-ordered)"
for w in 4 8; do $P 128 $w 2 0; done
echo "# 256-byte rows"
$P 256 4 2 0; $P 256 4 2 -1
} 2>&1 | tee gpurun_out/store_probe_r2.txt
