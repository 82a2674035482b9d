"""Where the end-to-end step of the headline workload spends its time (developer tool, GPU box)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from mjhmc_b200.samplers.hmc_state import HMCState

w = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
sampler, dist, X0, V0 = bench.make_sampler(w, 0)
Xh = torch.as_tensor(X0).pin_memory(); Vh = torch.as_tensor(V0).pin_memory()
iters = w["iters"]
def sync(): torch.cuda.synchronize()
def T(f, n=4):
    f(); sync(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    sync(); return (time.perf_counter() - t0) / n * 1e3
def step():
    sampler.state = HMCState.from_buffers(sampler, Xh, Vh)
    return sampler.sample(iters)
print("e2e step ms", T(step))
print("sample only ms", T(lambda: sampler.sample(iters)))
print("sample_device ms", T(lambda: sampler.sample_device(iters)))
S = sampler.sample_device(iters); sync()
out = torch.empty(S.shape, dtype=S.dtype, pin_memory=True)
print("one D2H of S ms (%.0f MB)" % (S.numel() * 8 / 1e6), T(lambda: out.copy_(S, non_blocking=True)))
def alloc():
    o = torch.empty(S.shape, dtype=S.dtype, pin_memory=True); return None
print("pinned alloc (cached) ms", T(alloc))
keep = []
def alloc2():
    keep.append(torch.empty(S.shape, dtype=S.dtype, pin_memory=True)); 
    if len(keep) > 2: keep.pop(0)
print("pinned alloc, two kept alive ms", T(alloc2))
def upload():
    sampler.state = HMCState.from_buffers(sampler, Xh, Vh); sampler.sample_device(1)
print("upload + 1 iteration ms", T(upload))
for ch in (1, 2, 4, 8, 16):
    type(sampler)._pipeline_chunks = lambda self, n, ch=ch: ch
    print("chunks", ch, "sample ms", T(lambda: sampler.sample(iters)))
