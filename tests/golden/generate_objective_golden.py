"""Golden vectors for the search objective's curve functions from the UNMODIFIED reference source.

    python tests/golden/generate_objective_golden.py      (build container only: needs /root/reference)

mjhmc/search/objective.py imports tensorflow and has Python-2 print statements, so it cannot be imported; the source
text of the pure functions ``curve_fn`` (:222-223), ``estimate_params`` (:190-218) and ``fit`` (:120-135) is read
from the reference and executed as it stands (numpy + scipy.optimize.curve_fit, both present).  Only inputs and
outputs are stored in objective_reference.npz.
"""
import os
import re

import numpy as np
from scipy.optimize import curve_fit

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/mjhmc/search/objective.py"


def _function_source(text, name):
    return re.search(r"^def %s\(.*?(?=^def |\Z)" % name, text, flags=re.S | re.M).group(0)


def main():
    text = open(SRC).read()
    ns = {"np": np, "curve_fit": curve_fit}
    for name in ("curve_fn", "estimate_params", "fit"):
        exec(compile(_function_source(text, name), SRC + ":" + name, "exec"), ns)
    rs = np.random.RandomState(3)
    out = {}
    cases = {"decay": (-1.7, 0.0), "osc": (-0.9, 6.0), "slow": (-0.2, 2.5)}
    for tag, (a, b) in cases.items():
        t = np.linspace(0, 2, 120)
        y = np.exp(a * t) * np.cos(b * t) + 0.01 * rs.randn(t.size)
        y[0] = 1.0
        out["t_" + tag] = t
        out["y_" + tag] = y
        out["curve_" + tag] = ns["curve_fn"](t, a, b)
        out["est_" + tag] = np.array(ns["estimate_params"](t, y), dtype=np.float64)
        out["fit_" + tag] = np.array(ns["fit"](t, y), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "objective_reference.npz"), **out)
    print({k: (v.tolist() if v.size <= 2 else v.shape) for k, v in out.items() if not k.startswith(("t_", "y_", "curve_"))})


if __name__ == "__main__":
    main()
