#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -2
for i in 1 2; do python bench.py --no-cpu-baseline --no-secondary --steps 10 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=j['e2e']
print('value %.4g e2e %.4g pcie %.1f share %.3f roof %.4g'%(j['value'], e['value'], e['pcie_gbs'], e['copy_time_share'], e['roof_value']))"; done
