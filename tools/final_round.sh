#!/bin/bash
# Record run on one B200: GPU tests, smoke(), the default bench line and the reference arm.  Output under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/final_gputest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/final_smoke.txt
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/final_bench_n1.err
timeout 600 python bench.py --config cfg1 --no-extras > gpurun_out/final_bench_cfg1.json 2>/dev/null; echo "cfg1 rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("final_bench_n1.json", "final_bench_cfg1.json", "final_bench_reference.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "clocks")}, (d.get("timing") or {}).get("build_us_median"), (d.get("timing") or {}).get("lookup_us_median"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
