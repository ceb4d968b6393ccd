#!/bin/bash
# Other bench modes after the bench.py edits + memcheck of the lookup kernels.  Output: gpurun_out/misc_*.{json,txt}
timeout 600 python bench.py --config cfg1 > gpurun_out/misc_cfg1.json 2> gpurun_out/misc_cfg1.err; echo "cfg1 rc=$?"
timeout 600 python bench.py --alternate > gpurun_out/misc_alt.json 2> gpurun_out/misc_alt.err; echo "alt rc=$?"
timeout 600 python bench.py --config cfg3 --no-extras > gpurun_out/misc_cfg3.json 2> gpurun_out/misc_cfg3.err; echo "cfg3 rc=$?"
python - <<'PY'
import json
for f in ("misc_cfg1.json", "misc_alt.json", "misc_cfg3.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "clocks")}, (d.get("roofline") or {}).get("frac"), (d.get("timing") or {}).get("burst"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lookup and not largest and not full" 2>&1 | tail -5 | tee gpurun_out/misc_memcheck.txt
