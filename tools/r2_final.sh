#!/bin/bash
# GPU box: the numbers and ncu evidence committed under profiles/ at the end of round 2.  usage: tools/r2_final.sh <tag>
tag=${1:-r2f}
o=gpurun_out
python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1; tail -2 $o/${tag}_smoke.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee $o/${tag}_pytest.log
python bench.py > $o/${tag}_bench_default.json 2> $o/${tag}_bench_default.err; cut -c1-400 $o/${tag}_bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_reference.json 2>/dev/null; cut -c1-300 $o/${tag}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -s 2 -c 1 -f -o $o/${tag}_prof_pot100d_mjhmc_f32 \
    python bench.py --workload pot100d_mjhmc_f32 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_pot_f32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_sample_kernel -s 2 -c 1 -f -o $o/${tag}_prof_pot100d_mjhmc \
    python bench.py --workload pot100d_mjhmc --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_pot_f64.log 2>&1
ls -la $o/${tag}_*
