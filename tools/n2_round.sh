#!/bin/bash
# N-rank sanity of the bench under torchrun (output: gpurun_out/nN_*.json).  usage: tools/n2_round.sh N
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench rc=$?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/n{n}_bench.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "clocks")}, (d.get("e2e") or {}).get("value"), d.get("cuda_graph_step"), (d.get("e2e_frames") or {}).get("value"))
except Exception as e:
    print("unreadable:", e)
PY
tail -c 300 gpurun_out/n${N}_bench.err
