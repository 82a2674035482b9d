#!/bin/bash
# ncu captures: Funnel register kernel after the shared refresh, the 2-d streaming points
o=gpurun_out
tag=${1:-r2n}
ncu --set full --clock-control none --import-source on -k regex:fused_sample -s 3 -c 1 -f -o $o/${tag}_prof_funnel10d_cthmc \
    python bench.py --workload funnel10d_cthmc --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_funnel.log 2>&1
for w in roughwell2d_control_L1_stream testgauss2d_control_L1_stream; do
  ncu --set full --clock-control none --import-source on -k regex:stream_sample -s 3 -c 1 -f -o $o/${tag}_prof_$w \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_$w.log 2>&1
done
ls -la $o | tail -5
