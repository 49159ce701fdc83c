#!/bin/bash
# dev: row-tiled tests (incl. the multi-level variant)
tag=${1:-s25}
timeout 400 python -m pytest tests/test_rowtiled.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_tests.log
