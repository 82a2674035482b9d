#!/bin/bash
# usage: tools/variant_many.sh "<lib tags: main or _variants names>" workload...   (GPU box)
tags=$1; shift
for t in $tags; do
  if [ "$t" = main ]; then unset MJHMC_B200_LIB; else export MJHMC_B200_LIB=$PWD/mjhmc_b200/_variants/lib_$t.so; fi
  for w in "$@"; do
    python bench.py --workload "$w" --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    try: j=json.loads(ln)
    except Exception: continue
    print('%-6s %-34s value %.4g  ms/step %.4f  %s frac %.3f' % ('$t', j['config']['workload'].split(':')[0], j['value'], j['ms_per_step'], j['roofline']['bound'], j['roofline']['frac']))
"
  done
done
