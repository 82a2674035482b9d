"""Debug: RoughWell-2d MJHMC at the searched hyper-parameters -- all statistics, GPU fp64 / fp32 vs oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import test_gpu_baseline_hp as t

name = sys.argv[1] if len(sys.argv) > 1 else "roughwell2d_mjhmc"
N, T = 6000, 60
o, orate = t._oracle_side(name, 3000, T)
res = {}
for dt in ("float64", "float32"):
    g, grate, s = t._gpu_side(name, N, T, dt)
    res[dt] = (g, grate)
print("%-9s %12s | %12s %6s | %12s %6s" % ("stat", "oracle", "gpu64", "z", "gpu32", "z"))
for k in sorted(o):
    row = "%-9s %12.5g |" % (k, o[k].mean())
    for dt in ("float64", "float32"):
        a = res[dt][0][k]
        se = np.sqrt(a.var(ddof=1) / len(a) + o[k].var(ddof=1) / len(o[k]))
        row += " %12.5g %6.2f |" % (a.mean(), abs(a.mean() - o[k].mean()) / se)
    print(row)
print("dEdX rate", orate, res["float64"][1], res["float32"][1])
