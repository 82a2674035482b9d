"""Fair initialisations by burn-in, and their cache (reference: mjhmc/misc/gen_mj_init.py and
Distribution.cached_init_X / load_cache, misc/distributions.py:104-149,182-195).

The reference burns every new distribution in for 1e6 sampling steps of 1000 particles with MarkovJumpHMC and
again with ControlHMC, estimating the variance of all sample scalars with a Python-level Welford loop --
about 2e9 particle-iterations of the hot path, hours to days on numpy.  Here the burn-in is a handful of fused
launches and the variance comes from device-side chunk moments merged with Chan's formula.

The cache is a pickle of the reference's 4-tuple ``(mjhmc_endpt, emc_var_estimate, true_var_estimate,
control_endpt)`` named ``<Class>_<hash>.pickle``.  Python 3 randomises ``hash()`` of strings per process, so the
file key is ``stable_hash(distribution)``: the distribution's own ``__hash__`` where that only hashes numbers,
and a CRC of the name for LambdaDistribution.
"""
import os
import pickle
import zlib

import numpy as np
import torch

from .. import _device, _lib

BURN_IN_STEPS = int(1E6)
VAR_STEPS = int(5E5)
MAX_N_PARTICLES = 1000

INIT_DIR = os.environ.get("MJHMC_B200_INIT_DIR") or os.path.join(
    os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "initializations")


def stable_hash(distribution):
    name = getattr(distribution, "name", None)
    if isinstance(name, str):
        return hash((distribution.ndims, zlib.crc32(name.encode())))
    return hash(distribution)


def cache_path(distribution):
    return os.path.join(INIT_DIR, '{}_{}.pickle'.format(type(distribution).__name__, stable_hash(distribution)))


def online_variance(sampler, distribution, var_steps=None, chunk=4096):
    """Variance of every sample scalar over `var_steps` sampling steps (gen_mj_init.py:76-98), unbiased
    (n - 1) like the Welford loop of the reference.  Returns (variance, sampler)."""
    var_steps = VAR_STEPS if var_steps is None else var_steps
    lib = _lib.load()
    count, mean, m2 = 0, 0.0, 0.0
    done = 0
    while done < var_steps:
        m = min(chunk, var_steps - done)
        S = sampler.sample_device(m)
        out = torch.zeros(2, dtype=torch.float64, device=S.device)
        _lib.check(lib.mjhmc_moments(_device.dtype_code(S.dtype), _device.ptr(S), S.numel(), _device.ptr(out),
                                     _device.stream_ptr(S.device)), "moments")
        s, s2 = out.cpu().tolist()
        nb = S.numel()
        mean_b = s / nb
        m2_b = s2 - s * mean_b
        # Chan et al. pairwise merge of (count, mean, M2)
        delta = mean_b - mean
        tot = count + nb
        mean += delta * nb / tot
        m2 += m2_b + delta * delta * count * nb / tot
        count = tot
        done += m
    return m2 / float(count - 1), sampler


def generate_initialization(distribution, burn_in_steps=None, var_steps=None, **sampler_kwargs):
    """gen_mj_init.py:14-52: burn MarkovJumpHMC in, estimate the variance of the embedded chain, keep its end
    point; the same for ControlHMC from a fresh gen_init_X."""
    from ..samplers.markov_jump_hmc import ControlHMC, MarkovJumpHMC
    burn_in_steps = BURN_IN_STEPS if burn_in_steps is None else burn_in_steps
    var_steps = VAR_STEPS if var_steps is None else var_steps
    print('Generating fair initialization for {} by burning in {} steps'.format(
        type(distribution).__name__, burn_in_steps))
    assert burn_in_steps > var_steps
    mjhmc = MarkovJumpHMC(distribution=distribution, resample=False, **sampler_kwargs)
    mjhmc._advance(burn_in_steps - var_steps, record=False)
    assert mjhmc.resample is False
    emc_var_estimate, mjhmc = online_variance(mjhmc, distribution, var_steps)
    # we discard v since p(x,v) = p(x)p(v)
    mjhmc_endpt = mjhmc.state.copy().X

    # otherwise will go into recursive loop
    distribution.mjhmc = False
    try:
        distribution.gen_init_X()
    except NotImplementedError:
        print("No explicit init method found, using mjhmc endpoint")
    distribution.E_count = 0
    distribution.dEdX_count = 0

    control = ControlHMC(distribution=distribution, **sampler_kwargs)
    control._advance(burn_in_steps - var_steps, record=False)
    true_var_estimate, control = online_variance(control, distribution, var_steps)
    control_endpt = control.state.copy().X
    return mjhmc_endpt, emc_var_estimate, true_var_estimate, control_endpt


def cache_initialization(distribution, **kwargs):
    """gen_mj_init.py:54-73: generate and pickle the fair initialisation of `distribution`
    (which must have nbatch == MAX_N_PARTICLES and generation_instance == True)."""
    distr_name = type(distribution).__name__
    result = generate_initialization(distribution, **kwargs)
    os.makedirs(INIT_DIR, exist_ok=True)
    path = cache_path(distribution)
    with open(path, 'wb') as cache_file:
        pickle.dump(result, cache_file)
    print("Fair initialization for {} saved as {}".format(distr_name, os.path.basename(path)))
    print("The embedded jump process on {} has estimated variance of {}".format(distr_name, result[1]))
    print("Meanwhile {} itself has an estimated variance of {}".format(distr_name, result[2]))
    return path
