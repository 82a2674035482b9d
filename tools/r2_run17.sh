#!/bin/bash
# 8-GPU box: the default bench line at N = 8 (weak legs, strong leg, secondaries) and at N = 8 with --scaling strong
o=gpurun_out
tag=${1:-r2w}
nvidia-smi topo -m 2>/dev/null | head -12 > $o/${tag}_topo.txt; nproc >> $o/${tag}_topo.txt; numactl -H 2>/dev/null | head -4 >> $o/${tag}_topo.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 \
   > $o/${tag}_bench_8gpu.json 2> $o/${tag}_bench_8gpu.err; tail -3 $o/${tag}_bench_8gpu.err
python - $o/${tag}_bench_8gpu.json <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N=%d main value %.4g e2e %.4g" % (j["n_gpus"], j["value"], j["e2e"]["value"]), j["e2e"], j.get("strong_scaling"))
for k,v in j.get("workloads",{}).items():
    if "error" in v: print(k, v); continue
    print("%-32s value %.4g ms/step %.3f" % (k, v["value"], v["ms_per_step"]), (v.get("ess") or {}).get("ess_per_s"))
PY
