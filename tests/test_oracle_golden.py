"""CPU: the oracle restatement against fixtures produced by the unmodified reference."""
import numpy as np
import pytest

from oracle import mjhmc_oracle as orc
from oracle import philox
from tests import helpers


@pytest.mark.parametrize("name", helpers.golden_inject_cases())
def test_injected_trajectory_bit_exact(name):
    g = helpers.load_inject(name)
    s = helpers.oracle_from_golden(name, g)
    n_iter = g["X"].shape[0]
    for it in range(n_iter):
        s.sampling_iteration()
        # same numpy operations in the same order -> identical bits
        np.testing.assert_array_equal(s.X, g["X"][it])
        np.testing.assert_array_equal(s.V, g["V"][it])
        np.testing.assert_array_equal(s.EX, g["EX"][it])
        np.testing.assert_array_equal(s.EV, g["EV"][it])
        c = s.counters()
        assert [c["l"], c["f"], c["fl"], c["r"], c["E"], c["dEdX"]] == list(g["counters"][it])
        assert s.attempt == int(g["attempts"][it])
        if name.startswith(("MarkovJumpHMC", "ContinuousTimeHMC")):
            np.testing.assert_array_equal(s.dwelling_times, g["dwell"][it])
        if name.startswith("MarkovJumpHMC"):
            np.testing.assert_array_equal(s.cache_active, g["cache"][it])
    assert s.epsilon == float(g["final_epsilon"]) and s.num_leapfrog_steps == int(g["final_L"])


def _build_seeded(run):
    """Replays the reference's constructor draw order."""
    np.random.seed(run["seed"])
    dk = run["distribution_kwargs"]
    d, N = dk["ndims"], dk["nbatch"]
    name = run["distribution"]
    if name == "RoughWell":
        energy = orc.RoughWellEnergy()
        gen = lambda: 100 * np.random.randn(d, N)
    elif name == "Gaussian":
        energy = orc.GaussianEnergy.log_conditioned(d, dk["log_conditioning"])
        cond = 10 ** np.linspace(-dk["log_conditioning"], 0, d)
        gen = lambda: (1. / np.sqrt(cond).reshape((-1, 1))) * np.random.randn(d, N)
    else:
        energy = orc.TestGaussianEnergy()
        gen = lambda: np.random.randn(d, N)
    gen()                 # Distribution() constructor
    X0 = gen()            # sampler constructor: distribution.reset()
    if run["sampler"] == "ContinuousTimeHMC":
        np.random.randn(d, N)   # first HMCState's V (state is built twice, Q18)
        X0 = gen()
    hp = run["hp"]
    return orc.OracleSampler(run["sampler"], energy, X0, epsilon=hp["epsilon"], beta=hp["beta"],
                             num_leapfrog_steps=hp["num_leapfrog_steps"],
                             resample=run["sampler_kwargs"].get("resample", True))


@pytest.mark.parametrize("idx", range(len(helpers.load_seeded()["runs"])))
def test_seeded_known_answers(idx):
    run = helpers.load_seeded()["runs"][idx]
    s = _build_seeded(run)
    X = s.sample(run["n_samples"])
    assert list(X.shape) == run["shape"]
    assert s.counters() == run["counters"]
    assert X.sum() == run["sum_X"]
    assert (X ** 2).sum() == run["sum_X2"]
    assert s.H().sum() == run["sum_H"]
    assert s.X.sum() == run["sum_final_X"] and s.V.sum() == run["sum_final_V"]


def test_backoff_known_answer():
    ka = helpers.load_seeded()["backoff"]
    np.random.seed(3)
    np.random.randn(1, 4); X0 = np.random.randn(1, 4)
    s = orc.OracleSampler("MarkovJumpHMC", orc.TestGaussianEnergy(), X0, epsilon=1.0, beta=0.5,
                          num_leapfrog_steps=1, resample=False)
    s.X[:] = [[100., .1, .2, .3]]
    s.V[:] = 0.
    s.EX = s.energy.E(s.X); s.EV = s._kinetic(s.V); s.g = s.energy.dEdX(s.X)
    e0, g0 = s.E_count, s.dEdX_count
    s.sampling_iteration()
    assert (s.E_count - e0, s.dEdX_count - g0) == (ka["dE"], ka["ddEdX"]) == (16, 24)
    assert (s.l_count, s.f_count, s.r_count) == (ka["l"], ka["f"], ka["r"])
    assert s.X.tolist() == ka["X"]
    assert (s.epsilon, s.num_leapfrog_steps) == (ka["epsilon"], ka["L"]) == (1.0, 1)
    assert list(s.cache_active) == ka["cache_active"]
    assert ka["ct_raises"]
    c = orc.OracleSampler("ContinuousTimeHMC", orc.TestGaussianEnergy(), s.X * 0 + [[100., .1, .2, .3]],
                          V=np.zeros((1, 4)), epsilon=1.0, beta=0.5, num_leapfrog_steps=1)
    with pytest.raises(ValueError):
        c.sampling_iteration()


def test_min_idx_semantics():
    """tests/test_utils.py:15-53 of the reference: column-wise argmin -> index sets."""
    rs = np.random.RandomState(1)
    a, b, c = rs.randn(3, 100)
    choice = np.argmin(np.stack([a, b, c]), axis=0)
    assert set(np.where(choice == 0)[0]) == set(np.arange(100)[(a < b) & (a < c)])
    assert set(np.where(choice == 1)[0]) == set(np.arange(100)[(b < a) & (b < c)])
    assert set(np.where(choice == 2)[0]) == set(np.arange(100)[(c < a) & (c < b)])


def test_resample_indices_match_reference_loop():
    rs = np.random.RandomState(5)
    dwell = rs.exponential(size=300)
    r = np.sort(rs.random_sample(300)) * dwell.sum()
    cumul = np.cumsum(dwell)
    expect = [np.where(cumul > v)[0][0] for v in r]      # markov_jump_hmc.py:326-328
    assert list(orc.resample_indices(dwell, r)) == expect


@pytest.mark.parametrize("energy,d", [
    (orc.ProductOfTEnergy(np.random.RandomState(0).randn(6, 6) * 0.5, np.random.RandomState(1).rand(6) * 2 + 2.1,
                          np.random.RandomState(2).randn(6) * 0.1), 6),
    (orc.FunnelEnergy(scale=3.0), 5),
    (orc.FunnelEnergy(scale=1.5, literal=True), 5),
    (orc.RoughWellEnergy(100, 4), 3),
    (orc.GaussianEnergy(np.random.RandomState(3).randn(4, 4)), 4),
    (orc.MultimodalGaussianEnergy(1, 3), 3),
])
def test_gradients_match_finite_differences(energy, d):
    rs = np.random.RandomState(7)
    X = rs.randn(d, 9)
    g = energy.dEdX(X)
    h = 1e-6
    for k in range(d):
        Xp, Xm = X.copy(), X.copy()
        Xp[k] += h
        Xm[k] -= h
        fd = (energy.E(Xp) - energy.E(Xm)) / (2 * h)
        np.testing.assert_allclose(g[k], fd, rtol=2e-6, atol=2e-7)


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), 0, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, 0xffffffffffffffff, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), 0x299f31d0a4093822,
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, out in kat:
        got = philox.philox4x32_10(*[np.array([c]) for c in ctr], key)
        assert tuple(int(w[0]) for w in got) == out


def test_philox_draw_statistics():
    n = 200000
    p = np.arange(n)
    u = philox.uniforms(123, 4, p)
    assert u.shape == (3, n) and u.min() >= 0 and u.max() < 1
    assert abs(u.mean() - 0.5) < 5e-3
    z = philox.normals(123, 4, p, 3)
    assert abs(z.mean()) < 1e-2 and abs(z.std() - 1) < 1e-2
    assert abs(np.corrcoef(z[0], z[1])[0, 1]) < 1e-2


def test_fft_autocor_is_circular_mean_product():
    rs = np.random.RandomState(0)
    s = rs.randn(2, 3, 16)
    ac = orc.fft_autocor(s)
    brute = np.array([np.mean(s * np.roll(s, -t, axis=-1)) for t in range(16)])
    np.testing.assert_allclose(ac, brute / brute[0], atol=1e-12)
    assert orc.ess_from_autocor(np.array([1.0, 0.5, 0.25, -0.1, 0.3])) == 5 / (1 + 2 * 0.75)


def test_oracle_fft_autocor_matches_the_reference_function():
    """autocor_reference.npz: outputs of the reference's own fft_autocor / slow_autocorrelation source
    (tests/golden/generate_autocor_golden.py), numpy.fft standing in for mklfft."""
    import os
    g = np.load(os.path.join(helpers.GOLDEN, "autocor_reference.npz"))
    for tag in "abc":
        x = g["x_" + tag]
        np.testing.assert_allclose(orc.fft_autocor(x), g["fft_" + tag], rtol=1e-12, atol=1e-13)
        T = x.shape[2]
        slow = np.array([np.mean(x ** 2)] + [np.mean(x[:, :, :-t] * x[:, :, t:]) for t in range(1, T - 1)])
        np.testing.assert_allclose(slow / slow[0], g["slow_" + tag], rtol=1e-12, atol=1e-13)
