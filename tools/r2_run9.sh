#!/bin/bash
# screened race: its own tests, the parity suites, the affected bench points
o=gpurun_out
tag=${1:-r2o}
timeout 900 python -m pytest tests/test_gpu_screen.py -q -x --timeout 300 2>&1 | tail -15
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_edges.py tests/test_gpu_stream.py -q -x --timeout 300 2>&1 | tail -6
tools/bench_many.sh $o/${tag}_lines.jsonl funnel10d_cthmc roughwell2d_mjhmc gauss100d_diag_mjhmc gauss100d_mjhmc_f32 pot100d_mjhmc_f32 funnel10d_cthmc_ess 2>&1 | grep -v "^$"
