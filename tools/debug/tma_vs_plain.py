"""Developer tool: where do the TMA-box and the plain-load/store paths of the tcgen05 kernel differ?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import helpers
from tests.test_gpu_parity import _dense_case, _counters
from mjhmc_b200 import _lib
from mjhmc_b200.samplers import markov_jump_hmc as S
lib = _lib.load()
d, N = 100, 60000
hp = dict(epsilon=0.12, beta=0.4, num_leapfrog_steps=2)
_, _, X0 = _dense_case("Gaussian", d, N, np.random.RandomState(9))
V0 = np.random.RandomState(10).randn(d, N)
out = []
for tma in (1, 0):
    dist, _, _ = _dense_case("Gaussian", d, N, np.random.RandomState(9))
    helpers.pin_init(dist, X0)
    lib.mjhmc_stream_set_tma(tma)
    s = S.MarkovJumpHMC(distribution=dist, V=V0, seed=21, dtype="float32", resample=False, **hp)
    rec = {}
    rec["X1"] = s.sample(3, preserve_order=True); rec["sx1"] = s.state.X.copy(); rec["sv1"] = s.state.V.copy()
    rec["c1"] = np.array(_counters(s, dist)); rec["dw1"] = np.asarray(s.dwelling_times).copy()
    eng = s._engine
    for nm in ("Hc", "ca"):
        try:
            rec[nm] = getattr(eng, nm)[eng.cur].cpu().numpy().copy() if hasattr(eng, nm) else None
        except Exception as e:
            rec[nm] = None
    rec["X2"] = s.sample(2, preserve_order=True); rec["sx2"] = s.state.X.copy(); rec["sv2"] = s.state.V.copy()
    rec["c2"] = np.array(_counters(s, dist))
    out.append(rec)
lib.mjhmc_stream_set_tma(1)
a, b = out
for k in a:
    if a[k] is None:
        print(k, "n/a"); continue
    x, y = np.asarray(a[k]), np.asarray(b[k])
    bad = x != y
    print(k, x.shape, "mismatch", int(bad.sum()))
    if bad.any() and x.ndim >= 2:
        cols = np.where(bad.reshape(x.shape[0], -1).any(axis=0))[0] if x.ndim == 2 else np.where(bad.any(axis=(0, 2)))[0]
        print("   particles:", cols[:40], "count", len(cols))
        r = N * np.arange(149) // 148 // 4 * 4
        rel = [(int(c - r[np.searchsorted(r, c, side='right') - 1]), int(r[np.searchsorted(r, c, side='right')] - c)) for c in cols[:40]]
        print("   (offset from CTA range start, distance to range end):", rel)
        if x.ndim == 3:
            print("   iterations with mismatch:", np.where(bad.any(axis=(0, 1)))[0])
