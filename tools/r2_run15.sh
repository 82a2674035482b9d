#!/bin/bash
o=gpurun_out
tag=${1:-r2t}
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15
tools/bench_many.sh $o/${tag}_lines.jsonl roughwell2d_mjhmc roughwell2d_control roughwell10d_control_L1_stream gauss100d_mjhmc pot100d_mjhmc 2>&1 | grep -v "^$"
