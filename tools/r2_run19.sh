#!/bin/bash
o=gpurun_out
tag=${1:-r2y}
for w in pot100d_mjhmc_f32 gauss100d_mjhmc_f32; do
ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -s 3 -c 1 -f -o $o/${tag}_prof_$w \
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_$w.log 2>&1
done
ls -la $o/${tag}_prof_*
