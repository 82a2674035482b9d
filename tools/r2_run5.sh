#!/bin/bash
o=gpurun_out
tag=${1:-r2e}
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout 60 -k "dense_float32" 2>&1 | tail -5
for w in gauss100d_mjhmc_f32 pot100d_mjhmc_f32; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_bench_$w.json 2> $o/${tag}_bench_$w.err
  python - $o/${tag}_bench_$w.json <<'PY'
import json,sys
j=json.load(open(sys.argv[1])); r=j["roofline"]
print(j["config"]["workload"][:30], "value %.4g ms/step %.3f frac %.3f achieved %.4g peak %.4g e2e %.3g" % (j["value"], j["ms_per_step"], r["frac"], r["achieved"], r["peak"], j["e2e"]["value"]))
PY
done
if [ "$2" = ncu ]; then
for w in gauss100d_mjhmc_f32 pot100d_mjhmc_f32; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -s 3 -c 1 -f -o $o/${tag}_prof_$w \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_$w.log 2>&1
done
fi
