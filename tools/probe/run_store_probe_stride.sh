#!/bin/bash
P=tools/probe/tma_store_probe3
mkdir -p gpurun_out
{
for row in 256 128; do
for s in 28672 28800 28928 29184 29696 30720 32768 36864 28736; do $P $s 8 $row; done; done
for s in 28672 28928 32768; do $P $s 16 256; $P $s 8 512; done
} 2>&1 | tee gpurun_out/store_probe3_stride.txt
