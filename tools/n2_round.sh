#!/bin/bash
# N=2 sanity of the bench modes under torchrun (output: gpurun_out/n2_*.json)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench rc=$?"
timeout 600 $TR bench.py --gpus 2 --train --steps 8 > gpurun_out/n2_train.json 2> gpurun_out/n2_train.err; echo "train rc=$?"
timeout 300 $TR bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/n2_ref.json 2> gpurun_out/n2_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("n2_bench.json", "n2_train.json", "n2_ref.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "clocks")}, (d.get("e2e") or {}).get("value"), (d.get("timing") or {}).get("burst"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -c 400 gpurun_out/n2_bench.err gpurun_out/n2_train.err
