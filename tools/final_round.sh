#!/bin/bash
# Record run on one B200: GPU tests, smoke(), the default bench line and the reference arm.  Output under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/final_gputest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/final_smoke.txt
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/final_bench_n1.err
RCB_LIB_VARIANT=lpq4 timeout 900 python bench.py --no-extras > gpurun_out/final_bench_n1_lpq4.json 2>/dev/null; echo "lpq4 rc=$?"
timeout 900 python bench.py --no-extras > gpurun_out/final_bench_n1_noextras.json 2>/dev/null; echo "noextras rc=$?"
python - <<'PY'
import json
for f in ("final_bench_n1.json", "final_bench_n1_lpq4.json", "final_bench_n1_noextras.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "clocks")}, d["timing"]["build_us_median"], d["timing"]["lookup_us_median"], (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__cycles_active.avg,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:lookup_tma -c 4 --csv --log-file gpurun_out/lookup_lpq2_counts.csv python tools/time_lookup.py --reps 2 > gpurun_out/ncu_lookup.log 2>&1; echo "ncu rc=$?"
fi
