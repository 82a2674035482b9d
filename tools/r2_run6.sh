#!/bin/bash
# full GPU test suite + default bench (with secondaries)
o=gpurun_out
tag=${1:-r2j}
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25 > $o/${tag}_pytest.log; tail -8 $o/${tag}_pytest.log
timeout 900 python bench.py > $o/${tag}_bench_default.json 2> $o/${tag}_bench_default.err; tail -3 $o/${tag}_bench_default.err
python - $o/${tag}_bench_default.json <<'PY'
import json,sys
j=json.load(open(sys.argv[1]))
print("main value %.4g e2e %.4g frac %.3f" % (j["value"], j["e2e"]["value"], j["roofline"]["frac"]))
for k,v in j.get("workloads",{}).items():
    if "error" in v: print(k, v); continue
    r=v["roofline"]; e=v.get("ess") or {}
    print("%-32s value %.4g ms/step %.3f %s frac %.3f" % (k, v["value"], v["ms_per_step"], r["bound"], r["frac"]), {kk: e[kk] for kk in ("ess_per_chain","ess_per_s","sampling_ms","autocorr_ms","first_negative_lag") if kk in e})
PY
