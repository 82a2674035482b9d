#!/bin/bash
python bench.py --no-cpu-baseline > gpurun_out/r2l_bench_default.json 2> gpurun_out/r2l_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench_default.json'))
print('value %.4g e2e %.4g share %.3f'%(d['value'], d['e2e']['value'], d['e2e']['copy_time_share']))
for k,v in d.get('workloads',{}).items():
    rr=v.get('roofline',{})
    print("%-34s %.4g  %.3f ms  %s %.3f steps %s"%(k, v.get('value',0), v.get('ms_per_step',0), rr.get('bound'), rr.get('frac',0), v.get('steps')))
PY
