#!/bin/bash
# Round-2 final campaign: -m gpu suite, smoke(), default bench, bench --impl reference, ncu launch list, ncu full capture of k_lnconv.
tag=r02g
timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/${tag}_S_bench_default_invocation.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_S_bench_default_invocation.json')); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['cpu_baseline']['value'], d.get('parity_db'))"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_S_bf16.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --no-graph > gpurun_out/${tag}_ncu.log 2>&1; echo ncu rc=$?; wc -l gpurun_out/${tag}_launches_S_bf16.csv
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:k_im2col_tc --launch-skip 8 -c 8 -o gpurun_out/${tag}_im2col python tools/profile_step.py > gpurun_out/${tag}_ncu_full.log 2>&1; echo ncu full rc=$?; ls -la gpurun_out/${tag}_im2col.ncu-rep
timeout 120 python tools/launch_table.py --min-ms 0 > gpurun_out/${tag}_launch_table_S_bf16.txt 2>&1; head -1 gpurun_out/${tag}_launch_table_S_bf16.txt
