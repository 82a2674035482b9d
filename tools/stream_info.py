"""Developer tool (GPU box): launch geometry the streaming kernel picks for a few shapes."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mjhmc_b200 import _lib
from mjhmc_b200.misc import distributions as D
from mjhmc_b200.samplers.markov_jump_hmc import ControlHMC

lib = _lib.load()
for d, N, dtype in [(2, 1 << 20, "float64"), (10, 1 << 20, "float64"), (16, 1 << 19, "float64"), (16, 1 << 19, "float32"),
                    (100, 1 << 17, "float64"), (128, 1 << 16, "float32"), (6, 1001, "float64")]:
    dist = D.TestGaussian(d, N)
    s = ControlHMC(distribution=dist, epsilon=0.5, beta=0.1, num_leapfrog_steps=1, dtype=dtype, kernel="stream", seed=1)
    s.sampling_iteration()
    out = (C.c_int64 * 7)()
    lib.mjhmc_stream_last_launch(out)
    print("d=%d N=%d %s: tma=%d stages=%d grid=%d ctas/sm=%d G=%d DT=%d smem=%d" % ((d, N, dtype) + tuple(out)), flush=True)
