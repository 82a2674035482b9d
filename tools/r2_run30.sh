#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "dense or pot or ProductOfT or Gaussian or gauss" 2>&1 | tail -2
tools/variant_many.sh "main hint st6 st8 main" testgauss2d_control_L1_stream roughwell2d_control_L1_stream
tools/variant_many.sh "main" pot100d_mjhmc gauss100d_mjhmc
ncu --set full --clock-control none --import-source on -k regex:dense_sample_kernel -s 2 -c 1 -f -o gpurun_out/r2g_prof_pot100d_mjhmc \
    python bench.py --workload pot100d_mjhmc --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2g_ncu_pot_f64.log 2>&1
