#!/bin/bash
# row-tiled RawFormer-L on N GPUs for every N given: bash tools/rt_multi.sh <tag> N [N...]
tag=$1; shift
for n in "$@"; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --row-tiled --size L --steps 20 > gpurun_out/${tag}_L_n$n.json 2> gpurun_out/${tag}_L_n$n.err; echo rc=$?
  tail -1 gpurun_out/${tag}_L_n$n.err
  python - <<PY
import json
for l in open("gpurun_out/${tag}_L_n$n.json"):
    if l.startswith("{"):
        d = json.loads(l); print($n, round(d["value"], 1), round(d["ms_per_step"], 3), d["parity"], d["sync_ms_per_step"], d["e2e"]["value"])
PY
done
