#!/bin/bash
# 2-GPU box: NCCL tests of the sharded couplings + the default bench line at N = 2 (weak legs, strong leg, secondaries)
o=gpurun_out
tag=${1:-r2v}
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 600 2>&1 | tail -8 > $o/${tag}_multi.log; tail -4 $o/${tag}_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
   > $o/${tag}_bench_2gpu.json 2> $o/${tag}_bench_2gpu.err; tail -3 $o/${tag}_bench_2gpu.err
python - $o/${tag}_bench_2gpu.json <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N=%d main value %.4g e2e %.4g" % (j["n_gpus"], j["value"], j["e2e"]["value"]), j.get("strong_scaling"))
for k,v in j.get("workloads",{}).items():
    if "error" in v: print(k, v); continue
    print("%-32s value %.4g ms/step %.3f" % (k, v["value"], v["ms_per_step"]), (v.get("ess") or {}).get("ess_per_s"))
PY
