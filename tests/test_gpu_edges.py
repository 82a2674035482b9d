"""GPU: edge cases of the hot path -- every register-kernel dimension incl. padded ones, tiny and ragged
particle counts, zero rates (infinite holding times), zero iterations, large strides."""
import numpy as np
import pytest

from oracle import mjhmc_oracle as orc
from tests import helpers

pytestmark = pytest.mark.gpu


def _counters(sampler, dist):
    return [sampler.l_count, sampler.f_count, sampler.fl_count, sampler.r_count, dist.E_count, dist.dEdX_count]


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 16])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_every_register_kernel_dimension(d, dtype):
    """ndims 1..16 map to the template dims {1,2,3,4,6,8,10,16}; the padded rows must not leak into energies."""
    from mjhmc_b200.misc.distributions import RoughWell
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    rs = np.random.RandomState(d)
    N = 130                                  # one full CTA + 2 particles
    X0, V0 = rs.randn(d, N) * 2, rs.randn(d, N)
    dist = helpers.pin_init(RoughWell(d, N, scale1=4, scale2=3), X0)
    hp = dict(epsilon=0.2, beta=0.3, num_leapfrog_steps=2)
    s = MarkovJumpHMC(distribution=dist, V=V0, seed=5, dtype=dtype, resample=False, **hp)
    assert s._engine.fused
    o = orc.OracleSampler("MarkovJumpHMC", orc.RoughWellEnergy(4, 3), X0, V=V0, draws=orc.PhiloxDraws(5),
                          resample=False, **hp)
    n = 2
    X, Xo = s.sample(n), o.sample(n)
    if dtype == "float64":
        assert helpers.rel_err(X, Xo) < 1e-10
        assert _counters(s, dist) == list(o.counters()[k] for k in ("l", "f", "fl", "r", "E", "dEdX"))
    else:
        same = np.all(np.abs(X - Xo) <= 1e-3 * (1 + np.abs(Xo)), axis=0)
        assert same.mean() > 0.97 and helpers.rel_err32(X[:, same], Xo[:, same]) < 1e-4


@pytest.mark.parametrize("N", [1, 2, 31, 33, 127, 129])
@pytest.mark.parametrize("kind", ["ControlHMC", "ContinuousTimeHMC", "MarkovJumpHMC"])
def test_ragged_particle_counts(kind, N):
    from mjhmc_b200.misc.distributions import Funnel
    from mjhmc_b200.samplers import markov_jump_hmc as S
    rs = np.random.RandomState(N)
    d = 5
    X0, V0 = rs.randn(d, N) * 0.5, rs.randn(d, N)
    dist = helpers.pin_init(Funnel(scale=2.0, nbatch=N, ndims=d), X0)
    hp = dict(epsilon=0.1, beta=0.4, num_leapfrog_steps=3)
    extra = dict(resample=False) if kind != "ControlHMC" else {}
    s = getattr(S, kind)(distribution=dist, V=V0, seed=9, **hp, **extra)
    o = orc.OracleSampler(kind, orc.FunnelEnergy(2.0), X0, V=V0, draws=orc.PhiloxDraws(9), resample=False, **hp)
    X, Xo = s.sample(4, preserve_order=True), o.sample(4, preserve_order=True)
    assert X.shape == (d, N, 4)
    assert helpers.rel_err(X, Xo) < 1e-10
    assert _counters(s, dist) == list(o.counters()[k] for k in ("l", "f", "fl", "r", "E", "dEdX"))


def test_zero_rates_give_infinite_holding_times():
    """p_r = 0 (beta -> 0: README defaults) and an exactly reversible proposal (f rate 0): the zero-rate
    branches of draw_from (utils.py:38-40) -- dwelling times are +inf for those operators, L always wins."""
    from mjhmc_b200.misc.distributions import TestGaussian
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    rs = np.random.RandomState(0)
    X0, V0 = rs.randn(2, 50), rs.randn(2, 50)
    dist = helpers.pin_init(TestGaussian(2, 50), X0)
    s = MarkovJumpHMC(distribution=dist, V=V0, seed=1)           # epsilon=1e-4, L=5, alpha=.2 -> beta = 0.2**2000 = 0
    assert s.beta == 1 and s.p_r == 0.0
    o = orc.OracleSampler("MarkovJumpHMC", orc.TestGaussianEnergy(), X0, V=V0, draws=orc.PhiloxDraws(1), resample=False)
    s.resample = False
    X, Xo = s.sample(10), o.sample(10)
    assert helpers.rel_err(X, Xo) < 1e-12
    assert (s.l_count, s.f_count, s.r_count) == (o.l_count, o.f_count, o.r_count)
    assert s.r_count == 0
    assert np.all(np.isfinite(s.dwelling_times))


def test_sample_zero_and_one():
    from mjhmc_b200.misc.distributions import RoughWell
    from mjhmc_b200.samplers.markov_jump_hmc import ControlHMC
    rs = np.random.RandomState(2)
    X0 = rs.randn(2, 20)
    dist = helpers.pin_init(RoughWell(2, 20), X0)
    s = ControlHMC(distribution=dist, epsilon=0.5, beta=0.3, V=rs.randn(2, 20), seed=4)
    assert s.sample(0).shape == (2, 0)
    np.testing.assert_array_equal(s.state.X, X0)
    assert s.sample(1).shape == (2, 20) and s.sample(1, preserve_order=True).shape == (2, 20, 1)
    assert dist.E_count == 20 + 2 * 20


def test_state_assignment_and_hyperparameter_change_between_calls():
    """Callers of the reference edit sampler.state and the hyper-parameters between iterations
    (figures/poe_fig.py:59-74)."""
    from mjhmc_b200.misc.distributions import RoughWell
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    rs = np.random.RandomState(3)
    X0, V0 = rs.randn(2, 64) * 3, rs.randn(2, 64)
    dist = helpers.pin_init(RoughWell(2, 64, scale1=5, scale2=4), X0)
    hp = dict(epsilon=0.3, beta=0.3, num_leapfrog_steps=3)
    s = MarkovJumpHMC(distribution=dist, V=V0, seed=8, resample=False, **hp)
    o = orc.OracleSampler("MarkovJumpHMC", orc.RoughWellEnergy(5, 4), X0, V=V0, draws=orc.PhiloxDraws(8),
                          resample=False, **hp)
    s.sample(3); o.sample(3)
    # edit the state through the host view, change epsilon / L
    st = s.state
    st.X[:, :5] = 0.25
    st.V[:, 7] *= -1
    o.X[:, :5] = 0.25
    o.V[:, 7] *= -1
    o.EX = o.energy.E(o.X); o.EV = o._kinetic(o.V); o.g = o.energy.dEdX(o.X)
    # an edited host state comes back with the reference cache flags but conservatively without F-move energies
    s.epsilon = o.epsilon = 0.15
    s.num_leapfrog_steps = o.num_leapfrog_steps = 5
    X, Xo = s.sample(4), o.sample(4)
    assert helpers.rel_err(X, Xo) < 1e-10
    assert (s.l_count, s.f_count, s.r_count) == (o.l_count, o.f_count, o.r_count)
    assert helpers.rel_err(s.state.H(), (o.EX + o.EV).reshape(1, -1)) < 1e-10
