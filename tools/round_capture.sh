#!/bin/bash
# GPU box: the numbers and ncu evidence committed under profiles/ for this round.  usage: tools/round_capture.sh <tag>
tag=${1:-r1c}
o=gpurun_out
python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1; tail -2 $o/${tag}_smoke.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py > $o/${tag}_bench_default.json 2> $o/${tag}_bench_default.err; cut -c1-300 $o/${tag}_bench_default.json
if [ "$2" = ref ]; then python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_reference.json 2>/dev/null; fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_launches.log 2>&1
for w in gauss10d_control_L1_stream gauss100d_diag_control_L1 gauss100d_diag_mjhmc; do
  ncu --set full --clock-control none --import-source on -k regex:stream_sample -s 3 -c 1 -f -o $o/${tag}_prof_$w \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_$w.log 2>&1
done
tools/bench_many.sh $o/${tag}_bench_lines.jsonl roughwell2d_control funnel10d_cthmc testgauss2d_control_L1_stream roughwell2d_control_L1_stream \
    roughwell10d_control_L1_stream gauss10d_control_L1_stream gauss16d_control_L1_stream gauss16d_control_L1_f32_stream gauss100d_diag_control_L1 \
    gauss100d_diag_mjhmc gauss100d_diag_mjhmc_x8 funnel10d_cthmc_ess
