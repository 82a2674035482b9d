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


# ---------------------------------------------------------------------------------------------------------------
# N1: fair-initialisation burn-in against the reference's literal loop run on the oracle
# ---------------------------------------------------------------------------------------------------------------
def _oracle_online_variance(o, var_steps):
    """misc/gen_mj_init.py:76-98, literally: Welford over every scalar of sample(1), step by step."""
    curr_mean, curr_sumsq, trial_idx = 0.0, 0.0, 0
    for _ in range(var_steps):
        for val in o.sample(1).ravel():
            trial_idx += 1
            delta = val - curr_mean
            curr_mean += float(delta) / trial_idx
            curr_sumsq += delta * (val - curr_mean)
    return curr_sumsq / float(var_steps * o.nbatch * o.ndims - 1)


def test_generate_initialization_matches_the_reference_loop_on_the_oracle():
    """gen_mj_init.generate_initialization (burn-in of MarkovJumpHMC and ControlHMC, online variance, end points) on
    the GPU against misc/gen_mj_init.py:14-52 executed literally with the oracle samplers on the same Philox stream:
    300 burn-in steps, 100 of them feeding the variance.  (A quadratic energy: on the RoughWell a 1e-13 difference
    grows past the tolerance within 300 iterations.)"""
    from mjhmc_b200.misc import distributions as D, gen_mj_init
    rs = np.random.RandomState(3)
    d, N = 2, 64
    X0, V0 = rs.randn(d, N) * 3, rs.randn(d, N)
    hp = dict(epsilon=0.4, beta=0.3, num_leapfrog_steps=4)
    dist = helpers.pin_init(D.Gaussian(d, N, log_conditioning=1), X0)
    dist.generation_instance = True
    got = gen_mj_init.generate_initialization(dist, burn_in_steps=300, var_steps=100, seed=21, V=V0, **hp)

    energy = orc.GaussianEnergy.log_conditioned(d, 1)
    mj = orc.OracleSampler("MarkovJumpHMC", energy, X0, V=V0, draws=orc.PhiloxDraws(21), resample=False, **hp)
    for _ in range(200):
        mj.sampling_iteration()
    emc_var = _oracle_online_variance(mj, 100)
    ctl = orc.OracleSampler("ControlHMC", energy, X0, V=V0, draws=orc.PhiloxDraws(21), **hp)
    for _ in range(200):
        ctl.sampling_iteration()
    true_var = _oracle_online_variance(ctl, 100)

    assert helpers.rel_err(got[0], mj.X) < 1e-9
    assert helpers.rel_err(got[3], ctl.X) < 1e-9
    np.testing.assert_allclose(got[1], emc_var, rtol=1e-9)
    np.testing.assert_allclose(got[2], true_var, rtol=1e-9)


# ---------------------------------------------------------------------------------------------------------------
# N4: state-ladder extraction (experiments/spectral.py) against the reference's counter-polling loops on the oracle
# ---------------------------------------------------------------------------------------------------------------
def _oracle_ladder_heatmap(o, max_steps):
    """experiments/spectral.py:96-131 literally, with the dihedral walk of samplers/algebraic_hmc.py:485-519."""
    last_r = last_l = last_f = 0
    k1, k2 = 0, 0
    visits = {}
    for _ in range(max_steps):
        o.sampling_iteration()
        if o.r_count != last_r:
            last_r += 1
            k1, k2 = 0, 0
        elif o.l_count != last_l:
            last_l += 1
            k2 += -1 if k1 else 1
        elif o.f_count != last_f:
            last_f += 1
            k1 ^= 1
        visits[(k2, k1)] = visits.get((k2, k1), 0) + 1
    return visits


def _oracle_ladder_generator(o, max_steps):
    """experiments/spectral.py:163-213 literally (dict keyed by the signed ladder position for the MAX_ORDER / 2 array)."""
    last_r = last_l = last_f = 0
    k1, k2 = 0, 0
    lad = {0: float(np.squeeze(o.H()))}
    out = []
    for _ in range(max_steps):
        o.sampling_iteration()
        if o.r_count != last_r:
            last_r += 1
            fwd, bwd = [], []
            j = 0
            while j in lad and lad[j] != 0:
                fwd.append(lad[j]); j += 1
            j = -1
            while j in lad and lad[j] != 0:
                bwd.append(lad[j]); j -= 1
            out.append(np.array(bwd[::-1] + fwd))
            k1, k2 = 0, 0
            lad = {0: float(np.squeeze(o.H()))}
        elif o.l_count != last_l:
            last_l += 1
            k2 += -1 if k1 else 1
            lad[k2] = float(np.squeeze(o.H()))
        elif o.f_count != last_f:
            last_f += 1
            k1 ^= 1
    return out


def _ladder_setup(N):
    from mjhmc_b200.misc import distributions as D
    rs = np.random.RandomState(8)
    X0, V0 = rs.randn(2, N) * 2, rs.randn(2, N)
    # (a quadratic energy: on the RoughWell rounding differences grow to 1e-7 within the 400 iterations of these tests)
    dist = helpers.pin_init(D.Gaussian(2, N, log_conditioning=1), X0)
    return dist, orc.GaussianEnergy.log_conditioned(2, 1), X0, V0


def test_ladder_heatmap_and_generator_match_the_reference_loops():
    from mjhmc_b200.experiments import spectral
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    hp = dict(epsilon=0.5, num_leapfrog_steps=3, beta=0.2)
    steps = 400
    # one chain, as the reference requires
    dist, energy, X0, V0 = _ladder_setup(1)
    heat = spectral.ladder_heatmap(MarkovJumpHMC, dist, max_steps=steps, seed=4, V=V0, chunk=64, window=64, **hp)
    o = orc.OracleSampler("MarkovJumpHMC", energy, X0, V=V0, draws=orc.PhiloxDraws(4), resample=False, **hp)
    assert heat == _oracle_ladder_heatmap(o, steps)
    dist, energy, X0, V0 = _ladder_setup(1)
    ladders = list(spectral.ladder_generator(MarkovJumpHMC, dist, max_steps=steps, seed=4, V=V0, chunk=64, **hp))
    o = orc.OracleSampler("MarkovJumpHMC", energy, X0, V=V0, draws=orc.PhiloxDraws(4), resample=False, **hp)
    want = _oracle_ladder_generator(o, steps)
    assert len(ladders) == len(want) > 5
    for a, b in zip(ladders, want):
        np.testing.assert_allclose(a, b, rtol=1e-9)
    # several chains: the device walk pools the visits of independent chains
    N = 5
    dist, energy, X0, V0 = _ladder_setup(N)
    heat = spectral.ladder_heatmap(MarkovJumpHMC, dist, max_steps=steps, seed=4, V=V0, chunk=100, window=64, **hp)
    pooled = {}
    for i in range(N):
        o = orc.OracleSampler("MarkovJumpHMC", energy, X0[:, i:i + 1], V=V0[:, i:i + 1], draws=orc.PhiloxDraws(4, i),
                              resample=False, **hp)
        for key, cnt in _oracle_ladder_heatmap(o, steps).items():
            pooled[key] = pooled.get(key, 0) + cnt
    assert heat == pooled


def test_ladder_numerical_err_hist_matches_the_reference_loop():
    """experiments/spectral.py:14-49 (ControlHMC, runs between R events) on the GPU trace against the literal loop."""
    from mjhmc_b200.experiments import spectral
    dist, energy, X0, V0 = _ladder_setup(1)
    hp = dict(epsilon=0.5, num_leapfrog_steps=3, beta=0.3)
    steps = 300
    cen, lens = spectral.ladder_numerical_err_hist(dist, n_steps=steps, seed=6, V=V0, chunk=50, **hp)
    o = orc.OracleSampler("ControlHMC", energy, X0, V=V0, draws=orc.PhiloxDraws(6), **hp)
    energies, run_lengths, r_counts = [], [], [0]
    ladder = [float(np.squeeze(o.H()))]
    run = 0
    for _ in range(steps):
        if o.r_count == r_counts[-1]:
            run += 1
            ladder.append(float(np.squeeze(o.H())))
        else:
            run_lengths.append(run)
            run = 0
            energies.append(np.array(ladder))
            ladder = [float(np.squeeze(o.H()))]
        r_counts.append(o.r_count)
        o.sampling_iteration()
    want = []
    for lad in energies:
        want += list(lad - lad[0])
    assert lens == run_lengths and len(lens) > 3
    np.testing.assert_allclose(cen, want, rtol=1e-9, atol=1e-10)


# ---------------------------------------------------------------------------------------------------------------
# N4: SparseImageCode (misc/tf_distributions.py:204-284) through the unfused device path
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("literal", [False, True])
@pytest.mark.parametrize("cauchy", [True, False])
def test_sparse_image_code_energy_and_sampler(literal, cauchy):
    from mjhmc_b200.misc.tf_distributions import SparseImageCode
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    rs = np.random.RandomState(5)
    n_patches, n_coeffs, img, N = 3, 24, 16, 6
    data = dict(data=rs.randn(img, 8), basis=rs.randn(img, n_coeffs) / 4)
    dist = SparseImageCode(n_patches=n_patches, n_batches=N, cauchy=cauchy, data=data, literal_reference_graph=literal)
    assert dist.ndims == n_patches * n_coeffs and dist.Xinit.shape == (dist.ndims, N)
    energy = orc.SparseImageCodeEnergy(data["basis"], data["data"][:, :n_patches].T, 0.01, cauchy, literal)
    X = rs.randn(dist.ndims, N)
    e0, g0 = dist.E_count, dist.dEdX_count
    np.testing.assert_allclose(dist.E(X), energy.E(X), rtol=1e-12)
    np.testing.assert_allclose(dist.dEdX(X), energy.dEdX(X), rtol=1e-11, atol=1e-13)
    assert (dist.E_count - e0, dist.dEdX_count - g0) == (N, N)
    # the sampler runs it in callback mode: device state, leapfrog pieces and transition kernels, this energy in between
    X0, V0 = rs.randn(dist.ndims, N) * 0.5, rs.randn(dist.ndims, N)
    helpers.pin_init(dist, X0)
    hp = dict(epsilon=0.05, beta=0.2, num_leapfrog_steps=3)
    s = MarkovJumpHMC(distribution=dist, V=V0, seed=31, resample=False, **hp)
    assert not s._engine.fused
    o = orc.OracleSampler("MarkovJumpHMC", energy, X0, V=V0, draws=orc.PhiloxDraws(31), resample=False, **hp)
    Xs, Xo = s.sample(4), o.sample(4)
    assert helpers.rel_err(Xs, Xo) < 1e-9
    c = o.counters()
    assert [s.l_count, s.f_count, s.fl_count, s.r_count, dist.E_count, dist.dEdX_count] == \
        [c["l"], c["f"], c["fl"], c["r"], c["E"], c["dEdX"]]
    with pytest.raises(IOError):
        SparseImageCode(n_patches=2, n_batches=1)            # the reference's blobs are not shipped
    syn = SparseImageCode(n_patches=2, n_batches=1, n_basis=512, synthetic=True)
    assert syn.ndims == 2 * 512 and syn.basis.shape == (256, 512)


# ---------------------------------------------------------------------------------------------------------------
# HMCState host view: the operators of samplers/hmc_state.py:86-129 on a handed-out state
# ---------------------------------------------------------------------------------------------------------------
def test_hmc_state_host_operators():
    from mjhmc_b200.misc import distributions as D
    from mjhmc_b200.samplers.markov_jump_hmc import ControlHMC
    rs = np.random.RandomState(2)
    d, N = 3, 17
    X0, V0 = rs.randn(d, N) * 2, rs.randn(d, N)
    dist = helpers.pin_init(D.RoughWell(d, N, scale1=5, scale2=4), X0)
    s = ControlHMC(distribution=dist, V=V0, seed=1, epsilon=0.3, beta=0.2, num_leapfrog_steps=4)
    energy = orc.RoughWellEnergy(5, 4)
    st = s.state.copy()
    g0, e0 = dist.dEdX_count, dist.E_count
    st.L()
    # reference: L leapfrogs of V -= eps/2 g; X += eps V; g = dEdX(X); V -= eps/2 g, then EV, EX (hmc_state.py:86-100)
    X, V, g = X0.copy(), V0.copy(), energy.dEdX(X0)
    for _ in range(4):
        V = V - 0.15 * g
        X = X + 0.3 * V
        g = energy.dEdX(X)
        V = V - 0.15 * g
    assert helpers.rel_err(st.X, X) < 1e-10 and helpers.rel_err(st.V, V) < 1e-10
    assert (dist.dEdX_count - g0, dist.E_count - e0) == (4 * N, N)            # counted like the reference
    np.testing.assert_allclose(np.ravel(st.H()), np.ravel(energy.E(X)) + np.sum(V ** 2, axis=0) / 2., rtol=1e-10)
    st.F()
    np.testing.assert_allclose(st.V, -V, rtol=1e-10)
    st2 = s.state.copy().FLF()
    Xf, Vf, gf = X0.copy(), -V0.copy(), energy.dEdX(X0)
    for _ in range(4):
        Vf = Vf - 0.15 * gf
        Xf = Xf + 0.3 * Vf
        gf = energy.dEdX(Xf)
        Vf = Vf - 0.15 * gf
    assert helpers.rel_err(st2.X, Xf) < 1e-10 and helpers.rel_err(st2.V, -Vf) < 1e-10
    # R: V = V sqrt(1 - beta) + randn sqrt(beta) with the sampler's beta (1 for ControlHMC: a full refresh)
    np.random.seed(0)
    Z = np.random.randn(d, N)
    np.random.seed(0)
    st3 = s.state.copy().R()
    np.testing.assert_allclose(st3.V, Z, rtol=1e-12)
    # update(idx, other): scatter-copy of the selected columns (hmc_state.py:63-72)
    a, b = s.state.copy(), st
    idx = np.array([0, 5, 9])
    a.update(idx, b)
    np.testing.assert_array_equal(a.X[:, idx], b.X[:, idx])
    np.testing.assert_array_equal(a.X[:, 1], X0[:, 1])
    # an edited state assigned back is what the sampler continues from
    s.state = st
    np.testing.assert_allclose(s.state.X, st.X)
