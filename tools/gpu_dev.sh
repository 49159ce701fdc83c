#!/bin/bash
# dev: full GPU suite, then A/B (alternating) of a debugging-aid environment variable on the default bench
tag=${1:-s29}; var=${2:-RAWFORMER_B200_GUIDE_SIDE}
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_tests.log
timeout 60 python tools/launch_table.py --min-ms 0 2>&1 | grep "guidance\|launches" | head -8
for k in 1 2 3; do for v in 0 1; do
env $var=$v timeout 100 python bench.py --no-cpu --no-extra --steps 30 > gpurun_out/${tag}_bench_$v$k.json 2> gpurun_out/${tag}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_$v$k.json')); print('$var=$v', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'])"
done; done
