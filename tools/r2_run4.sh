#!/bin/bash
o=gpurun_out
for w in gauss100d_mjhmc_f32 pot100d_mjhmc_f32; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $o/r2d_bench_$w.json 2> $o/r2d_bench_$w.err
  python - $o/r2d_bench_$w.json <<'PY'
import json,sys
j=json.load(open(sys.argv[1])); r=j["roofline"]
print(j["config"]["workload"][:30], "value %.4g ms/step %.3f frac %.3f achieved %.4g peak %.4g e2e %.3g" % (j["value"], j["ms_per_step"], r["frac"], r["achieved"], r["peak"], j["e2e"]["value"]))
PY
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -s 3 -c 1 -f -o $o/r2d_prof_$w \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/r2d_ncu_$w.log 2>&1
done
ls -la $o | grep r2d
