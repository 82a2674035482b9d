"""Developer tool (GPU box): bandwidth of the streaming kernel's data path alone (n_iter = 0: X, V in -> out through
the TMA ring and the coalesced stores), i.e. the ceiling of the HBM-bound workloads for a given shape."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mjhmc_b200.misc import distributions as D
from mjhmc_b200.samplers.markov_jump_hmc import ControlHMC

from mjhmc_b200 import _lib
lib = _lib.load()
torch.zeros(1, device="cuda")
for nb in (1, 2, 3, 4, 5, 6, 8):
    lo, hi = 0, 232448
    while lo < hi:
        mid = (lo + hi + 1) // 2
        if lib.mjhmc_stream_probe_blocks(mid) >= nb:
            lo = mid
        else:
            hi = mid - 1
    print("largest dynamic smem with %d resident CTAs of 256 threads: %d bytes" % (nb, lo), flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for d, N, dtype in [(2, 16_000_000, "float64"), (10, 8_000_000, "float64"), (16, 4_000_000, "float64"),
                    (16, 8_000_000, "float32"), (100, 1_000_000, "float64")]:
    dist = D.TestGaussian(d, 8)
    dist.nbatch = N
    X0 = np.zeros((d, N))
    dist.gen_init_X = lambda: setattr(dist, "Xinit", X0)
    s = ControlHMC(distribution=dist, epsilon=0.5, beta=0.1, num_leapfrog_steps=1, dtype=dtype, kernel="stream", seed=1,
                   V=X0)
    eng = s._engine
    S = 8 if dtype == "float64" else 4
    ts = []
    for k in range(8):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.launch(0, 0)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = np.median(ts[3:])
    print("d=%d N=%d %s: copy path %.3f ms  %.0f GB/s (4 d S bytes per particle)" % (d, N, dtype, t, 4 * d * S * N / t / 1e6),
          flush=True)
    del s, eng
    torch.cuda.empty_cache()
