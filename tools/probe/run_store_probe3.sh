#!/bin/bash
# What bounds one warp's box-store iteration?  variants: 0 baseline, 1 no staging write + no fence, 2 fence only,
# 4 = coalesced st.global lines (transposed-epilogue emulation, no TMA)
P=tools/probe/tma_store_probe
mkdir -p gpurun_out
{
for v in 0 1 2; do for d in 1 2 8; do $P 128 4 $d -1 28672 0 2 $v; done; done
for w in 2 4 8 16; do $P 128 $w 2 -1 28672 0 2 0; done
for w in 8 16; do $P 256 $w 2 -1 28672 0 2 0; done
$P 256 4 8 -1 28672 0 2 1
echo "# coalesced st.global"
for w in 4 8 16; do $P 128 $w 2 -1 28672 0 2 4; done
for w in 4 8; do $P 128 $w 2 0 28672 0 2 4; done
} 2>&1 | tee gpurun_out/store_probe_variants_r2.txt
