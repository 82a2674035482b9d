"""Golden vectors for the autocorrelation row from the UNMODIFIED reference functions.

    python tests/golden/generate_autocor_golden.py        (build container only: needs /root/reference)

mjhmc/misc/autocor.py cannot be imported under Python 3 (print statements elsewhere in the module, and
``mklfft`` is absent), so the source text of the two pure functions ``fft_autocor`` (autocor.py:37-49) and
``slow_autocorrelation`` (:177-211) is read from the reference at generation time and executed as it stands,
with ``fftn`` / ``ifftn`` bound to numpy.fft (the same transform mklfft wraps).  ``slow_autocorrelation`` is only
run with half_window=False: its half-window branch relies on Python-2 integer division.  No reference source is
stored; only inputs and outputs go to autocor_reference.npz.
"""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/mjhmc/misc/autocor.py"


def _function_source(text, name):
    m = re.search(r"^def %s\(.*?(?=^def |\Z)" % name, text, flags=re.S | re.M)
    return m.group(0)


def main():
    text = open(SRC).read()
    ns = {"np": np, "fftn": np.fft.fftn, "ifftn": np.fft.ifftn}
    for name in ("fft_autocor", "slow_autocorrelation"):
        exec(compile(_function_source(text, name), SRC + ":" + name, "exec"), ns)
    rs = np.random.RandomState(7)
    out = {}
    for tag, (d, N, T) in {"a": (3, 7, 16), "b": (2, 33, 41), "c": (10, 5, 64)}.items():
        # an AR(1)-like series so the curve is not just noise
        x = np.zeros((d, N, T))
        x[:, :, 0] = rs.randn(d, N)
        for t in range(1, T):
            x[:, :, t] = 0.8 * x[:, :, t - 1] + 0.6 * rs.randn(d, N)
        e = np.arange(T, dtype=np.float64)
        out["x_" + tag] = x
        out["fft_" + tag] = ns["fft_autocor"](x)
        slow, _, _ = ns["slow_autocorrelation"](x, e, e, half_window=False)
        out["slow_" + tag] = slow
    np.savez_compressed(os.path.join(HERE, "autocor_reference.npz"), **out)
    print("wrote autocor_reference.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
