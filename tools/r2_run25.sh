#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "dense or float32 or f32 or tc or baseline or screen" 2>&1 | tail -5
tools/variant_many.sh "main main" gauss100d_mjhmc_f32 pot100d_mjhmc_f32
for w in gauss100d_mjhmc_f32 pot100d_mjhmc_f32; do
MJHMC_B200_LIB=$PWD/mjhmc_b200/_variants/lib_timing.so python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-secondary 2>/dev/null | grep "^tid" | tail -4 | tee gpurun_out/r2z_timing2_$w.txt
done
