#!/bin/bash
# ncu captures after the screened race: Funnel register kernel, 100-d diagonal Gaussian MJHMC stream kernel
o=gpurun_out
tag=${1:-r2p}
ncu --set full --clock-control none --import-source on -k regex:fused_sample -s 3 -c 1 -f -o $o/${tag}_prof_funnel10d_cthmc \
    python bench.py --workload funnel10d_cthmc --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_funnel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stream_sample -s 3 -c 1 -f -o $o/${tag}_prof_gauss100d_diag_mjhmc \
    python bench.py --workload gauss100d_diag_mjhmc --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_g100.log 2>&1
python tools/debug/e2e_breakdown.py > $o/${tag}_e2e_breakdown.log 2>&1
tail -20 $o/${tag}_e2e_breakdown.log
