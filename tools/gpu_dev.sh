#!/bin/bash
# full -m gpu suite + smoke + compute-sanitizer memcheck of the dense-conv forms on small frames
tag=${1:-s27}
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 500 compute-sanitizer --tool memcheck --print-limit 10 python tools/dev_lnconv.py > gpurun_out/${tag}_memcheck.log 2>&1; echo memcheck rc=$?; tail -6 gpurun_out/${tag}_memcheck.log
