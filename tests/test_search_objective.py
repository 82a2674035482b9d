"""Search objective: the curve functions against outputs of the reference's own source
(tests/golden/objective_reference.npz), the numpy Adam fit, and -- on the GPU -- the whole objective."""
import os

import numpy as np
import pytest

from tests import helpers

G = np.load(os.path.join(helpers.GOLDEN, "objective_reference.npz"))
CASES = ("decay", "osc", "slow")


@pytest.mark.parametrize("tag", CASES)
def test_curve_functions_match_the_reference(tag):
    from mjhmc_b200.search import objective as obj
    t, y = G["t_" + tag], G["y_" + tag]
    a, b = {"decay": (-1.7, 0.0), "osc": (-0.9, 6.0), "slow": (-0.2, 2.5)}[tag]
    np.testing.assert_array_equal(obj.curve_fn(t, a, b), G["curve_" + tag])
    np.testing.assert_array_equal(np.array(obj.estimate_params(t, y), dtype=np.float64), G["est_" + tag])
    np.testing.assert_allclose(np.array(obj.fit(t, y), dtype=np.float64), G["fit_" + tag], rtol=1e-9, atol=1e-12)


def test_fit_of_nan_curve_and_unpack():
    from mjhmc_b200.search import objective as obj
    assert tuple(obj.fit(np.arange(3.0), np.array([1.0, np.nan, 0.2]))) == (1E3, 0)          # objective.py:134-135
    assert obj.unpack_params({"epsilon": np.array([0.3]), "beta": [0.1], "L": 5}) == {"epsilon": 0.3, "beta": 0.1, "L": 5}


def test_adam_fit_descends_from_the_reference_initialisation():
    """tf_fit (objective.py:137-183) restated in numpy: never worse than its estimate_params start, and it
    recovers a clean exponential decay."""
    from mjhmc_b200.search import objective as obj
    for tag in CASES:
        t, y = G["t_" + tag], G["y_" + tag]
        a0, b0 = obj.estimate_params(t, y)
        a, b = obj.tf_fit(t, y, n_steps=3000)
        loss = lambda p, q: float(np.sum((y - obj.curve_fn(t, p, q)) ** 2))
        assert loss(a, b) <= loss(a0, b0) + 1e-12
    t = np.linspace(0, 2, 200)
    a, b = obj.tf_fit(t, np.exp(-1.3 * t), n_steps=4000)
    assert abs(a + 1.3) < 0.02 and abs(b) < 1e-6


def test_min_idx():
    import pandas as pd
    from mjhmc_b200.search import objective as obj
    df = pd.DataFrame({"num grad": [10, 20, 30, 40], "autocorrelation": [1.0, 0.6, 0.4, 0.1]})
    assert obj.min_idx(df.copy(), 0.5) == 30 and obj.min_idx(df.copy(), 0.01) is None


@pytest.mark.gpu
def test_objective_on_the_gpu_scores_a_faster_sampler_lower():
    """obj_func end to end: fused sampling launches, device autocorrelation, fit.  A reasonable step size must
    score a more negative decay rate than a tiny one on the same gradient budget."""
    from mjhmc_b200.misc.distributions import TestGaussian
    from mjhmc_b200.samplers.markov_jump_hmc import ControlHMC
    from mjhmc_b200.search import objective as obj
    scores = []
    for eps in (0.9, 0.02):
        np.random.seed(5)
        dist = TestGaussian(ndims=2, nbatch=400)
        cos_coef, t, exp_coef, ac, _ = obj.obj_func_helper(
            ControlHMC, dist, False, dict(epsilon=eps, beta=0.2, num_leapfrog_steps=3, seed=3),
            overrides=dict(num_grad_steps=3000, use_cached_var=False))
        assert np.isfinite(exp_coef) and np.isfinite(cos_coef) and len(t) == len(ac)
        scores.append(exp_coef)
    assert scores[0] < scores[1] and scores[0] < 0


@pytest.mark.gpu
def test_objective_with_its_default_cached_variance_flag_runs_without_a_cache_file(tmp_path, monkeypatch):
    """obj_func passes use_cached_var=True like the reference (search/objective.py:40-47); the fair-initialisation
    cache it would read is opt-in here, so without the file the call must warn and go on (the fft estimate
    never reads the cached variance, autocor.py:107-111) instead of raising FileNotFoundError."""
    from mjhmc_b200.misc import distributions as D
    from mjhmc_b200.samplers.markov_jump_hmc import ControlHMC
    from mjhmc_b200.search import objective as obj
    np.random.seed(6)
    dist = D.TestGaussian(ndims=2, nbatch=300)
    from mjhmc_b200.misc import gen_mj_init
    monkeypatch.setattr(gen_mj_init, "INIT_DIR", str(tmp_path))     # no cache file can exist here
    cos_coef, t, exp_coef, ac, _ = obj.obj_func_helper(
        ControlHMC, dist, False, dict(epsilon=0.8, beta=0.2, num_leapfrog_steps=3, seed=4),
        overrides=dict(num_grad_steps=2000))
    assert np.isfinite(exp_coef) and np.isfinite(cos_coef) and len(t) == len(ac)
