#!/bin/bash
timeout 600 python -m pytest tests/test_search_objective.py tests/test_gpu_api.py -m gpu -q --timeout 300 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2k_prof_gauss100d_mjhmc_f32 python bench.py --workload gauss100d_mjhmc_f32 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2k_ncu_gauss.log 2>&1; tail -1 gpurun_out/r2k_ncu_gauss.log
