#!/bin/bash
# ncu evidence for the round: launch list of a bench run + full captures of the two dominant kernels.
mkdir -p gpurun_out
python tools/prof_path.py --mode f16f8 --lookups 6 --builds 3 > gpurun_out/plain_prof_path.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:build_tc_kernel -s 1 -c 1 -o gpurun_out/r2_build \
    python tools/prof_path.py --mode f16f8 --lookups 6 --builds 3 > gpurun_out/ncu_build.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lookup_tma_kernel -s 2 -c 2 -o gpurun_out/r2_lookup \
    python tools/prof_path.py --mode f16f8 --lookups 6 --builds 3 > gpurun_out/ncu_lookup.log 2>&1
ncu --set full --clock-control none -k regex:pack_operands_kernel -s 1 -c 1 -o gpurun_out/r2_pack \
    python tools/prof_path.py --mode f16f8 --lookups 6 --builds 3 > gpurun_out/ncu_pack.log 2>&1
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep
