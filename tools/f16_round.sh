#!/bin/bash
out=gpurun_out/f16_lpq.txt
: > $out
timeout 600 python -m pytest tests -m gpu -q -x -k "lanes or fp16 or sweep or fast" 2>&1 | tail -3 >> $out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_f16lpq.json 2>/dev/null; echo "bench rc=$?" >> $out
python - >> $out <<'PY'
import json
d=json.loads(open("gpurun_out/bench_f16lpq.json").read().strip().splitlines()[-1])
print(d["value"], d["fast_mode"])
PY
cat $out
