"""Shared helpers for the parity tests (oracle side)."""
import glob
import json
import os

import numpy as np

from oracle import mjhmc_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_inject_cases():
    return sorted(os.path.basename(p)[len("inject_"):-4] for p in glob.glob(os.path.join(GOLDEN, "inject_*.npz")))


def load_inject(name):
    return dict(np.load(os.path.join(GOLDEN, "inject_%s.npz" % name)))


def load_seeded():
    with open(os.path.join(GOLDEN, "seeded_runs.json")) as f:
        return json.load(f)


def energy_from_golden(g):
    dist = str(g["dist"])
    p = g["dist_params"]
    if dist == "RoughWell":
        return orc.RoughWellEnergy(p[0], p[1])
    if dist == "TestGaussian":
        return orc.TestGaussianEnergy(p[0])
    if dist == "Gaussian":
        return orc.GaussianEnergy(p)
    raise KeyError(dist)


def oracle_from_golden(name, g, draws=None):
    kind = name.split("_")[0]
    if draws is None:
        draws = orc.InjectedDraws(g["Z"], g["U"], g["U0"])
    return orc.OracleSampler(kind, energy_from_golden(g), g["X0"], V=g["V0"], epsilon=float(g["epsilon"]),
                             beta=float(g["beta_arg"]), num_leapfrog_steps=int(g["L"]), draws=draws,
                             resample=False)
