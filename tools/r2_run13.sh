#!/bin/bash
o=gpurun_out
tag=${1:-r2r}
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -12 > $o/${tag}_pytest.log; tail -5 $o/${tag}_pytest.log
timeout 900 python bench.py > $o/${tag}_bench_default.json 2> $o/${tag}_bench_default.err; tail -3 $o/${tag}_bench_default.err
python - $o/${tag}_bench_default.json <<'PY'
import json,sys
j=json.load(open(sys.argv[1]))
print("main value %.4g e2e %.4g frac %.3f" % (j["value"], j["e2e"]["value"], j["roofline"]["frac"]), j["e2e"])
for k,v in j.get("workloads",{}).items():
    if "error" in v: print(k, v); continue
    r=v["roofline"]; e=v.get("ess") or {}
    print("%-32s value %.4g ms/step %.3f %s frac %.3f" % (k, v["value"], v["ms_per_step"], r["bound"], r["frac"]), {kk: e[kk] for kk in ("ess_per_chain","ess_per_s","sampling_ms","autocorr_ms","first_negative_lag") if kk in e})
PY
ncu --set full --clock-control none --import-source on -k regex:stream_sample -s 3 -c 1 -f -o $o/${tag}_prof_roughwell10d_control_L1_stream \
    python bench.py --workload roughwell10d_control_L1_stream --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_rw10.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_sample -s 3 -c 1 -f -o $o/${tag}_prof_roughwell2d_mjhmc \
    python bench.py --workload roughwell2d_mjhmc --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_rw2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_stash -s 3 -c 1 -f -o $o/${tag}_prof_funnel10d_cthmc \
    python bench.py --workload funnel10d_cthmc --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $o/${tag}_ncu_funnel.log 2>&1
