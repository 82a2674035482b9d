"""GPU: the 'next' rows of SURVEY 8(f) -- generate_samples / autocorrelation drivers (N2) and the
fair-initialisation burn-in with its cache (N1)."""
import os
import pickle

import numpy as np
import pytest

from oracle import mjhmc_oracle as orc
from tests import helpers

pytestmark = pytest.mark.gpu


def _oracle_generate_samples(o, num_steps=None, num_grad_steps=None):
    """misc/autocor.py:213-261 on the oracle sampler (resample=False)."""
    N = float(o.nbatch)
    num_steps = num_steps or int(num_grad_steps // o.num_leapfrog_steps) + 100
    o.E_count = o.dEdX_count = 0                       # distribution.reset()
    samples = np.zeros((o.ndims, o.nbatch, num_steps))
    g, e = np.zeros(num_steps), np.zeros(num_steps)
    for t in range(num_steps):
        samples[:, :, t] = o.sample(1)
        g[t], e[t] = o.dEdX_count / N, o.E_count / N
        if num_grad_steps is not None and g[t] >= num_grad_steps:
            return samples[:, :, :t + 1], e[:t + 1], g[:t + 1]
    if num_grad_steps is not None:
        sel = g <= num_grad_steps
        return samples[:, :, sel], e[sel], g[sel]
    return samples, e, g


@pytest.mark.parametrize("kind", ["ControlHMC", "MarkovJumpHMC"])
@pytest.mark.parametrize("budget", [dict(num_steps=37), dict(num_grad_steps=120)])
def test_generate_samples_matches_reference_loop(kind, budget):
    from mjhmc_b200.misc import autocor
    from mjhmc_b200.misc.distributions import RoughWell
    from mjhmc_b200.samplers import markov_jump_hmc as S
    rs = np.random.RandomState(3)
    d, N = 2, 60
    X0, V0 = rs.randn(d, N) * 3, rs.randn(d, N)
    dist = helpers.pin_init(RoughWell(d, N, scale1=5, scale2=4), X0)
    hp = dict(epsilon=0.5, beta=0.3, num_leapfrog_steps=4)
    extra = dict(resample=False) if kind == "MarkovJumpHMC" else {}
    samples, e_evals, grad_evals = autocor.generate_samples(getattr(S, kind), dist, V=V0, seed=11, chunk=16,
                                                            **budget, **hp, **extra)
    o = orc.OracleSampler(kind, orc.RoughWellEnergy(5, 4), X0, V=V0, draws=orc.PhiloxDraws(11), resample=False, **hp)
    so, eo, go = _oracle_generate_samples(o, **budget)
    assert samples.shape == so.shape
    np.testing.assert_array_equal(e_evals, eo)
    np.testing.assert_array_equal(grad_evals, go)
    # ~150 leapfrog steps of a chaotic rough-well trajectory: ulp-level differences have grown to ~1e-10
    assert helpers.rel_err(samples, so) < 1e-8
    assert (dist.E_count, dist.dEdX_count) == (o.E_count, o.dEdX_count)      # stops where the reference stops


def test_calculate_autocorrelation_and_brute_force():
    from mjhmc_b200.misc import autocor
    from mjhmc_b200.misc.distributions import Gaussian
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    rs = np.random.RandomState(2)
    dist = helpers.pin_init(Gaussian(ndims=3, nbatch=50, log_conditioning=1), rs.randn(3, 50))
    V0 = rs.randn(3, 50)
    ac, e_evals, g_evals = autocor.calculate_autocorrelation(MarkovJumpHMC, dist, num_steps=40, epsilon=0.5, beta=0.2,
                                                             num_leapfrog_steps=3, resample=False, seed=4, V=V0)
    assert ac.shape == e_evals.shape == g_evals.shape == (40,) and ac[0] == 1.0
    # the same run, samples on the host, against the numpy formulas of the reference
    samples, e2, g2 = autocor.generate_samples(MarkovJumpHMC, dist.reset(), num_steps=40, epsilon=0.5, beta=0.2,
                                               num_leapfrog_steps=3, resample=False, seed=4, V=V0)
    np.testing.assert_allclose(ac, orc.fft_autocor(samples), atol=1e-10)
    T = samples.shape[2]
    for half in (False, True):
        slow, _, _ = autocor.slow_autocorrelation(samples, e2, g2, half_window=half)
        n_l = (T // 2) - 1 if half else T - 1
        c = np.array([np.mean(samples ** 2)] + [np.mean(samples[:, :, :-t] * samples[:, :, t:]) for t in range(1, n_l)])
        np.testing.assert_allclose(slow, c / c[0], atol=1e-10)
        bf, eb, gb = autocor.autocorrelation(samples, e2, g2, half_window=half, brute_force=True)
        max_t = int(T / 2) - 1 if half else T - 1
        cc = np.array([np.mean(samples[:, :, :T - t] * samples[:, :, t:]) for t in range(max_t)])
        np.testing.assert_allclose(bf[:, 0], np.concatenate(([1.0], cc[1:] / cc[0])), atol=1e-10)
        assert len(eb) == len(gb) == (int(T / 2) - 1 if half else T - 1)


def test_fair_initialisation_cache(tmp_path, monkeypatch):
    from mjhmc_b200.misc import distributions as D
    from mjhmc_b200.misc import gen_mj_init
    from mjhmc_b200.samplers.markov_jump_hmc import ControlHMC, MarkovJumpHMC
    monkeypatch.setattr(gen_mj_init, "INIT_DIR", str(tmp_path))
    monkeypatch.setattr(gen_mj_init, "MAX_N_PARTICLES", 256)
    np.random.seed(6)
    dist = D.TestGaussian(ndims=2, nbatch=40)
    assert gen_mj_init.stable_hash(dist) == gen_mj_init.stable_hash(D.TestGaussian(ndims=2, nbatch=7))
    dist.mjhmc = True
    dist.cached_init_X(burn_in_steps=300, var_steps=200, epsilon=0.6, beta=0.3, num_leapfrog_steps=3, seed=5)
    path = gen_mj_init.cache_path(dist)
    assert os.path.exists(path) and dist.Xinit.shape == (2, 40) and dist.nbatch == 40 and not dist.generation_instance
    with open(path, "rb") as f:
        mj_end, emc_var, true_var, ctl_end = pickle.load(f)
    assert mj_end.shape == ctl_end.shape == (2, 256)
    assert abs(true_var - 1.0) < 0.1 and emc_var > 0.5                  # unit Gaussian
    np.testing.assert_array_equal(dist.Xinit, mj_end[:, :40])
    assert dist.load_cache()[1] == emc_var
    dist.mjhmc = False
    dist.init_X(fair_init=True)
    np.testing.assert_array_equal(dist.Xinit, ctl_end[:, :40])
    # online_variance == unbiased numpy variance of the same samples
    d2 = helpers.pin_init(D.TestGaussian(ndims=2, nbatch=30), np.random.RandomState(1).randn(2, 30))
    s = ControlHMC(distribution=d2, epsilon=0.6, beta=0.3, num_leapfrog_steps=3, seed=9, V=np.zeros((2, 30)))
    var, _ = gen_mj_init.online_variance(s, d2, var_steps=50, chunk=16)
    d3 = helpers.pin_init(D.TestGaussian(ndims=2, nbatch=30), np.random.RandomState(1).randn(2, 30))
    s3 = ControlHMC(distribution=d3, epsilon=0.6, beta=0.3, num_leapfrog_steps=3, seed=9, V=np.zeros((2, 30)))
    np.testing.assert_allclose(var, np.var(s3.sample(50), ddof=1), rtol=1e-12)


def test_autocorrelation_kernels_match_the_reference_functions():
    """The device kernels behind fft_autocor / slow_autocorrelation against outputs of the reference's own
    functions (tests/golden/autocor_reference.npz, generate_autocor_golden.py)."""
    import os
    from mjhmc_b200.misc import autocor
    g = np.load(os.path.join(helpers.GOLDEN, "autocor_reference.npz"))
    for tag in "abc":
        x = g["x_" + tag]
        np.testing.assert_allclose(autocor.fft_autocor(x), g["fft_" + tag], rtol=1e-10, atol=1e-12)
        e = np.arange(x.shape[2], dtype=np.float64)
        slow, _, _ = autocor.slow_autocorrelation(x, e, e, half_window=False)
        np.testing.assert_allclose(slow, g["slow_" + tag], rtol=1e-10, atol=1e-12)
        ac, _, _ = autocor.autocorrelation(x, e, e, half_window=False, brute_force=False)
        np.testing.assert_allclose(ac, g["fft_" + tag], rtol=1e-10, atol=1e-12)
