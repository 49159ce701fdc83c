#!/bin/bash
# One GPU-box call: full -m gpu suite, smoke(), default bench, per-launch table.  Outputs in gpurun_out/.
tag=${1:-chk}
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['cpu_baseline']['value'], d.get('parity_db'), d['clocks'])"
timeout 120 python tools/launch_table.py --min-ms 0 > gpurun_out/${tag}_launch_table_S_bf16.txt 2>&1; head -1 gpurun_out/${tag}_launch_table_S_bf16.txt
