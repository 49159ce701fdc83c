#!/bin/bash
# Final check of a round: -m gpu suite, smoke(), the default bench invocation, then the ncu launch list of the same command.
tag=${1:-final}
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['cpu_baseline']['value'])"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-graph > gpurun_out/${tag}_ncu.log 2>&1; echo ncu rc=$?; wc -l gpurun_out/${tag}_launches.csv
