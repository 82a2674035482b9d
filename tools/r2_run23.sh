#!/bin/bash
tools/variant_many.sh "main fix8 fix16" gauss100d_mjhmc pot100d_mjhmc
MJHMC_B200_LIB=$PWD/mjhmc_b200/_variants/lib_trace.so python bench.py --workload gauss100d_mjhmc_f32 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary 2>/dev/null | grep "^st" | head -12 > gpurun_out/r2z_trace_gauss.txt
MJHMC_B200_LIB=$PWD/mjhmc_b200/_variants/lib_trace.so python bench.py --workload pot100d_mjhmc_f32 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary 2>/dev/null | grep "^st" | head -12 > gpurun_out/r2z_trace_pot.txt
cat gpurun_out/r2z_trace_gauss.txt
