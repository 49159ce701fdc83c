#!/bin/bash
# dev: A/B (alternating) of a debugging-aid environment variable on the default bench
tag=${1:-s28}; var=${2:-RAWFORMER_B200_LNCONV_WPRE}
timeout 60 python tools/dev_lnconv.py > gpurun_out/${tag}_dev.log 2>&1; echo dev rc=$?; tail -3 gpurun_out/${tag}_dev.log
for k in 1 2 3; do for v in 0 1; do
env $var=$v timeout 100 python bench.py --no-cpu --no-extra --steps 30 > gpurun_out/${tag}_bench_$v$k.json 2> gpurun_out/${tag}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_$v$k.json')); print('$var=$v', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'])"
done; done
