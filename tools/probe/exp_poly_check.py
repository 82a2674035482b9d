"""Error of dists.cuh exp_poly against mpmath, every fp64 operation emulated with exact rounding (developer tool).
Prints the worst relative error in units of 2^-53."""
import mpmath as mp
import numpy as np

mp.mp.dps = 60
C = [1.0, 1.0, 0.5000000000000019, 0.1666666666666668, 0.04166666666648795, 0.008333333333319589,
     0.0013888888952352863, 0.00019841269890076403, 2.4801485441561313e-05, 2.755724088722987e-06,
     2.763265472252779e-07, 2.5110049204818658e-08]
LOG2E, MAGIC, LN2HI, LN2LO = 1.4426950408889634, 6755399441055744.0, 0.6931471805599453, 2.3190468138462996e-17


def fma(a, b, c):
    return float(mp.mpf(a) * mp.mpf(b) + mp.mpf(c))


rng = np.random.default_rng(0)
worst = 0.0
for x in np.concatenate([rng.uniform(-700, 700, 3000), rng.uniform(-30, 30, 6000), rng.uniform(-1, 1, 3000)]):
    x = float(x)
    t = fma(x, LOG2E, MAGIC)
    k = t - MAGIC
    r = fma(k, -LN2LO, fma(k, -LN2HI, x))
    s = C[11]
    for v in reversed(C[:11]):
        s = fma(s, r, v)
    err = abs(mp.mpf(s) * mp.mpf(2) ** int(k) / mp.exp(mp.mpf(x)) - 1)
    worst = max(worst, float(err))
print("worst relative error %.3g = %.2f x 2^-53" % (worst, worst / 2 ** -53))
