#!/bin/bash
# One GPU-box call: full -m gpu suite, default bench (PDL on / off), row-tiled single band.  Outputs in gpurun_out/.
tag=${1:-chk}
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_tests.log
timeout 300 python bench.py --torch-eager > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
RAWFORMER_B200_PDL=0 timeout 200 python bench.py --no-cpu > gpurun_out/${tag}_bench_nopdl.json 2> gpurun_out/${tag}_bench_nopdl.err
timeout 200 python bench.py --row-tiled --size L --steps 10 > gpurun_out/${tag}_rt_L_n1.json 2> gpurun_out/${tag}_rt_L_n1.err; tail -2 gpurun_out/${tag}_rt_L_n1.err
python - <<PY
import json
for f in ("${tag}_bench", "${tag}_bench_nopdl", "${tag}_rt_L_n1"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), "MP/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"],
              "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), d.get("parity"), d.get("eager_b200_baseline"))
    except Exception as e:
        print(f, "FAILED", e)
PY
