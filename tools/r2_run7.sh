#!/bin/bash
# register-kernel checks: parity tests + the elementwise bench points
o=gpurun_out
tag=${1:-r2m}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_edges.py -q -x --timeout 120 2>&1 | tail -4
tools/bench_many.sh $o/${tag}_lines.jsonl funnel10d_cthmc roughwell2d_mjhmc roughwell2d_control funnel10d_cthmc_ess 2>&1 | grep -v "^$" | sed 's/--no-secondary//'
