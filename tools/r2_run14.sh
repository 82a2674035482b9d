#!/bin/bash
o=gpurun_out
tag=${1:-r2s}
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -5
tools/bench_many.sh $o/${tag}_lines.jsonl roughwell2d_mjhmc roughwell2d_control roughwell2d_control_L1_stream roughwell10d_control_L1_stream testgauss2d_control_L1_stream funnel10d_cthmc 2>&1 | grep -v "^$"
