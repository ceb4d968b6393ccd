#!/bin/bash
# Where do the ~2 us per lookup between the tight C-ABI loop and the bench step go?  Output: gpurun_out/lookup_gap.txt
out=gpurun_out/lookup_gap.txt
: > $out
T="timeout 300 python tools/time_lookup.py --reps 128"
$T >> $out 2>&1
$T --alloc >> $out 2>&1
$T --via-block >> $out 2>&1
$T --via-block --no-grad >> $out 2>&1
$T --via-block --reps 32 >> $out 2>&1
$T --alloc --reps 512 >> $out 2>&1
PYTORCH_NO_CUDA_MEMORY_CACHING=0 PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True $T --alloc >> $out 2>&1
cat $out
