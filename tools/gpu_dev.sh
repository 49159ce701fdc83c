#!/bin/bash
# full -m gpu suite + smoke
tag=${1:-s24}
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
