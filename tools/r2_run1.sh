#!/bin/bash
# GPU box, round 2 first pass: full GPU tests, default bench line (with the secondary workloads), Funnel ncu capture.
tag=${1:-r2a}
o=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_baseline_hp.py 2>&1 | tail -15 > $o/${tag}_pytest.log; tail -5 $o/${tag}_pytest.log
timeout 900 python -m pytest tests/test_gpu_baseline_hp.py -q -s 2>&1 | tail -40 > $o/${tag}_pytest_hp.log; tail -25 $o/${tag}_pytest_hp.log
timeout 900 python bench.py > $o/${tag}_bench_default.json 2> $o/${tag}_bench_default.err; cut -c1-600 $o/${tag}_bench_default.json; tail -3 $o/${tag}_bench_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_sample_kernel -s 3 -c 1 -f -o $o/${tag}_prof_funnel \
    python bench.py --workload funnel10d_cthmc --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_funnel.log 2>&1
ls -la $o | tail -5
