#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -5
tools/variant_many.sh "main main" gauss100d_mjhmc_f32 pot100d_mjhmc_f32
for w in gauss100d_mjhmc_f32 pot100d_mjhmc_f32; do
MJHMC_B200_LIB=$PWD/mjhmc_b200/_variants/lib_timing.so python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-secondary 2>/dev/null | grep "^tid" | tail -4 | tee gpurun_out/r2z_timing3_$w.txt
done
