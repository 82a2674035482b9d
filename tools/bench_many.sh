#!/bin/bash
# usage: tools/bench_many.sh out.jsonl workload1 workload2 ...   (GPU box)
out=$1; shift
: > "$out"
for w in "$@"; do
  python bench.py --workload "$w" --steps 10 --warmup 3 --no-cpu-baseline --no-secondary >> "$out" 2>> "${out%.jsonl}.err" || echo "{\"workload\": \"$w\", \"failed\": true}" >> "$out"
done
python - "$out" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    try:
        j = json.loads(ln)
    except Exception:
        continue
    if "value" not in j:
        print(j); continue
    print("%-34s value %.4g  ms/step %.4f  %s frac %.3f  e2e %.3g" % (j["config"]["workload"].split(":")[0], j["value"],
          j["ms_per_step"], j["roofline"]["bound"], j["roofline"]["frac"], j["e2e"]["value"]))
PY
