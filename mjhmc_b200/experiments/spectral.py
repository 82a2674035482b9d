"""State-ladder extraction from a running sampler (reference: mjhmc/experiments/spectral.py).

The reference watches ONE chain (``nbatch == 1``) and polls ``sampler.r_count / l_count / f_count`` and
``sampler.state.H()`` after every ``sampling_iteration()`` -- a host round trip per step.  Here the fused kernels
record the operator every particle took (``choice``) and, for the ladder energies, ``H()`` after every iteration
(``mjhmc_outputs.energy``); the ladder walk of ``ladder_heatmap`` runs on the device for all particles at once
(``mjhmc_ladder_visits``), the energy ladders of ``ladder_generator`` are cut out of the trace on the host.
With ``nbatch == 1`` the results are the reference's; more particles are independent chains whose ladders are
pooled (heat map) or yielded chain after chain (generator).

Out of scope here, as in SURVEY section 2 (C9): the algebraic ladder samplers and the spectral-gap figure built on
top of these ladders (``sp_img_ladder_generator`` needs the absent image blobs, ``test_fig`` matplotlib).
"""
import numpy as np
import torch

from .. import _device, _lib
from ..misc.distributions import Distribution, Gaussian
from ..samplers.markov_jump_hmc import ControlHMC

MAX_ORDER = int(1e10)

# ladder codes of the walk kernel: 0 = L, 1 = F, 2 = R, 3 = no move on the ladder
_L, _F, _R, _NONE = 0, 1, 2, 3


def _ladder_codes(sampler, choice):
    """Operator choices recorded by the kernels -> ladder moves, following the counter tests of
    spectral.py:112-124 (R first, then L, then F; a bare FL move changes none of the three counters)."""
    code = sampler._sampler_code
    if code == _lib.SAMPLER_MARKOV_JUMP:                  # 0 = L, 1 = F, 2 = R already
        return choice
    if code == _lib.SAMPLER_CONTINUOUS_TIME:              # 0 = F, 1 = FL, 2 = R
        table = torch.tensor([_F, _NONE, _R], dtype=torch.uint8, device=choice.device)
        return table[choice.long()]
    # discrete: bit0 accepted, bit1 flipped, bit2 the batch-wide R coin (markov_jump_hmc.py:138-148)
    table = torch.tensor([_NONE, _NONE, _F, _L, _R, _R, _R, _R], dtype=torch.uint8, device=choice.device)
    return table[choice.long()]


def _make_sampler(sampler_class, distribution, epsilon, num_leapfrog_steps, beta, **kwargs):
    assert isinstance(distribution, Distribution)
    if sampler_class.__name__ in ("ContinuousTimeHMC", "MarkovJumpHMC"):
        kwargs.setdefault("resample", False)
    return sampler_class(distribution=distribution, epsilon=epsilon, num_leapfrog_steps=num_leapfrog_steps, beta=beta,
                         **kwargs)


def ladder_heatmap(sampler_class, distribution, epsilon, num_leapfrog_steps, beta, max_steps=int(1e5), window=512,
                   chunk=4096, **kwargs):
    """ Computes a heatmap over ladder indices (spectral.py:72-131)

    Args:
      sampler_class: sampler to use - sampler class, ie not an object, the initializer
      distribution: the distribution to test - Distribution object (the reference demands nbatch == 1; more particles
         are more chains, their visits are summed)
      epsilon, num_leapfrog_steps, beta: sampler hyper-parameters
      max_steps: number of sampling steps to run
      window: ladder positions -window .. window are resolved (B200: the table is dense, not a dict)

    Returns:
      {(k_idx, p_idx): visits} with k_idx the signed position on the ladder and p_idx the flip bit -- the format
      ``unwrap_heatmap`` (spectral.py:215-243) produces.  Visits outside the window are under the key ``"outside"``.
    """
    sampler = _make_sampler(sampler_class, distribution, epsilon, num_leapfrog_steps, beta, **kwargs)
    eng = sampler._engine
    lib = _lib.load()
    W = 2 * window + 1
    with eng.ctx():
        state = torch.zeros((sampler.nbatch, 2), dtype=torch.int32, device=eng.device)
        visits = torch.zeros(2 * W + 1, dtype=torch.int64, device=eng.device)
        done = 0
        while done < max_steps:
            m = min(chunk, max_steps - done)
            _, _, choice = sampler._advance(m, record=False, want_choice=True)
            codes = _ladder_codes(sampler, choice).contiguous()
            _lib.check(lib.mjhmc_ladder_visits(_device.ptr(codes), m, sampler.nbatch, window, _device.ptr(state),
                                               _device.ptr(visits), eng._stream()), "ladder_visits")
            done += m
        v = visits.cpu().numpy()
    out = {}
    for p_idx in (0, 1):
        row = v[p_idx * W:(p_idx + 1) * W]
        for j in np.nonzero(row)[0]:
            out[(int(j) - window, p_idx)] = int(row[j])
    if v[2 * W]:
        out["outside"] = int(v[2 * W])
    return out


def _trace(sampler, n_steps, chunk):
    """(ladder codes, H after every iteration) of n_steps iterations as host arrays (T, nbatch), chunk by chunk."""
    done = 0
    while done < n_steps:
        m = min(chunk, n_steps - done)
        _, _, choice, energy = sampler._advance(m, record=False, want_choice=True, want_energy=True)
        yield _ladder_codes(sampler, choice).cpu().numpy(), energy.cpu().numpy()
        done += m


def ladder_generator(sampler_class, distribution, epsilon=0.0001, num_leapfrog_steps=5, beta=0.3, max_steps=int(1e5),
                     chunk=4096, **kwargs):
    """ Returns a generator over the ladders encountered while sampling (spectral.py:137-213): every R move closes
    the current ladder and yields its energies [H(L^-b z), ..., H(z), ..., H(L^f z)] in ladder order.

    The reference keeps the energies in an array of MAX_ORDER / 2 slots indexed by the ladder position modulo
    MAX_ORDER / 2 and reads the runs of non-zero entries from both ends; here they live in a dict keyed by the signed
    position.  Energies exist for the register-resident kernels (ndims <= 16) and callable energies."""
    sampler = _make_sampler(sampler_class, distribution, epsilon, num_leapfrog_steps, beta, **kwargs)
    N = sampler.nbatch
    H0 = np.asarray(sampler.state.H()).reshape(-1)
    sampler._host_state = None
    ladders = [{0: float(H0[i])} for i in range(N)]
    pos = np.zeros((N, 2), dtype=np.int64)               # [k1, k2] per chain
    for codes, energy in _trace(sampler, max_steps, chunk):
        for t in range(codes.shape[0]):
            for i in range(N):
                c = codes[t, i]
                if c == _R:
                    lad = ladders[i]
                    forward, backward = [], []
                    j = 0
                    while j in lad and lad[j] != 0:
                        forward.append(lad[j]); j += 1
                    j = -1
                    while j in lad and lad[j] != 0:
                        backward.append(lad[j]); j -= 1
                    assert len(forward) + len(backward) < (MAX_ORDER / 2)
                    yield np.array(backward[::-1] + forward)
                    ladders[i] = {0: float(energy[t, i])}
                    pos[i] = 0
                elif c == _L:
                    pos[i, 1] += -1 if pos[i, 0] else 1
                    ladders[i][int(pos[i, 1])] = float(energy[t, i])
                elif c == _F:
                    pos[i, 0] ^= 1


def ladder_numerical_err_hist(distr=None, n_steps=int(1e5), chunk=4096, **kwargs):
    """ Compute a histogram of the numerical integration error on the state ladder (spectral.py:14-49): ControlHMC,
    energies of the states visited between two R events, centred on the first of each run.

    Returns:
     centered_energies: list of H - H(first state of the run)
     run_lengths: list of observed run lengths
    """
    distr = distr or Gaussian(nbatch=1)
    sampler = ControlHMC(distribution=distr, **kwargs)
    N = sampler.nbatch
    H_now = np.asarray(sampler.state.H()).reshape(-1).astype(np.float64)
    sampler._host_state = None
    energies = [[] for _ in range(N)]
    run_lengths = [[] for _ in range(N)]
    current = [[float(H_now[i])] for i in range(N)]
    run_length = np.zeros(N, dtype=np.int64)
    fired_before = np.zeros(N, dtype=bool)                 # did the previous iteration fire R
    for codes, energy in _trace(sampler, n_steps, chunk):
        for t in range(codes.shape[0]):
            for i in range(N):
                # the reference tests the counter BEFORE stepping: it sees the R of the previous iteration
                if not fired_before[i]:
                    run_length[i] += 1
                    current[i].append(float(H_now[i]))
                else:
                    run_lengths[i].append(int(run_length[i]))
                    run_length[i] = 0
                    energies[i].append(np.array(current[i]))
                    current[i] = [float(H_now[i])]
            fired_before = codes[t] == _R
            H_now = energy[t]
    centered, lengths = [], []
    for i in range(N):
        for lad in energies[i]:
            centered += list(lad - lad[0])
        lengths += run_lengths[i]
    return centered, lengths


def fit_inv_pdf(ladder_energies):
    """ Interpolant of the inverse cdf of the ladder energies (spectral.py:51-69): lets callers draw energies from
    the empirical distribution of ladder_numerical_err_hist."""
    from scipy.interpolate import UnivariateSpline
    hist, bin_edges = np.histogram(ladder_energies, bins='auto')
    mid = bin_edges[:-1] + np.diff(bin_edges) / 2
    first = -2 * mid[0] + mid[1]                         # the reference's extrapolated first abscissa (:63)
    cdf = np.concatenate([[0], np.cumsum(hist) / np.sum(hist)])
    return UnivariateSpline(cdf, np.concatenate([[first], mid]), bbox=[0, 1], k=1)
