#!/bin/bash
# A/B of a debugging-aid environment variable on the default bench, ALTERNATING on one box (the pool's boxes differ by +-2 %
# and the second bench of a call runs hotter than the first, so only alternating runs on one box compare):
#   gpurun -- 'bash tools/gpu_ab.sh <tag> RAWFORMER_B200_LNCONV_PROJ'
tag=${1:-ab}; var=${2:-RAWFORMER_B200_LNCONV_PROJ}
for k in 1 2 3; do for v in 0 1; do
env $var=$v timeout 100 python bench.py --no-cpu --no-extra --steps 30 > gpurun_out/${tag}_bench_$v$k.json 2> gpurun_out/${tag}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_$v$k.json')); print('$var=$v', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'])"
done; done
