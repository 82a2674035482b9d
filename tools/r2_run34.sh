#!/bin/bash
MJHMC_BENCH_DEBUG=1 python bench.py --no-cpu-baseline --steps 5 2>&1 >/dev/null | grep "bench debug" | grep -E "stream|diag_control"
for s in 5 10; do MJHMC_BENCH_DEBUG=1 python bench.py --workload roughwell10d_control_L1_stream --no-cpu-baseline --no-secondary --steps $s 2>&1 >/dev/null | grep "bench debug"; done
