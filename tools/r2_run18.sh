#!/bin/bash
o=gpurun_out
tag=${1:-r2x}
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -k "dense or float32 or f32 or pot or tc or ProductOfT or baseline or statistics" 2>&1 | tail -6
tools/bench_many.sh $o/${tag}_lines.jsonl gauss100d_mjhmc_f32 pot100d_mjhmc_f32 2>&1 | grep -v "^$"
