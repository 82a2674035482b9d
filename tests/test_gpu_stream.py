"""GPU: the streaming separable-energy kernel (csrc/stream_separable.cuh; TMA ring, several threads per particle)
against the oracle, against the register kernel, and against itself (TMA vs cooperative loader, one launch vs many)."""
import numpy as np
import pytest

from oracle import mjhmc_oracle as orc
from tests import helpers

pytestmark = pytest.mark.gpu

KEYS = ("l", "f", "fl", "r", "E", "dEdX")


def _counters(sampler, dist):
    return [sampler.l_count, sampler.f_count, sampler.fl_count, sampler.r_count, dist.E_count, dist.dEdX_count]


def _make(dist_name, d, N, rs):
    from mjhmc_b200.misc import distributions as D
    if dist_name == "RoughWell":
        return D.RoughWell(d, N, scale1=4, scale2=3), orc.RoughWellEnergy(4, 3), rs.randn(d, N) * 2
    if dist_name == "TestGaussian":
        return D.TestGaussian(d, N, sigma=1.3), orc.TestGaussianEnergy(1.3), rs.randn(d, N)
    dist = D.Gaussian(ndims=d, nbatch=N, log_conditioning=2)
    return dist, orc.GaussianEnergy.log_conditioned(d, 2), rs.randn(d, N) / np.sqrt(np.diag(dist.J))[:, None]


def _pair(kind, dist_name, d, N, seed, hp, dtype="float64", kernel="stream", n_off=0):
    from mjhmc_b200.samplers import markov_jump_hmc as S
    rs = np.random.RandomState(seed)
    dist, energy, X0 = _make(dist_name, d, N, rs)
    V0 = rs.randn(d, N)
    helpers.pin_init(dist, X0)
    extra = dict(resample=False) if kind in ("ContinuousTimeHMC", "MarkovJumpHMC") else {}
    s = getattr(S, kind)(distribution=dist, V=V0, seed=seed, dtype=dtype, kernel=kernel, **hp, **extra)
    o = orc.OracleSampler(kind, energy, X0, V=V0, draws=orc.PhiloxDraws(seed), resample=False, **hp)
    return s, dist, o


@pytest.mark.parametrize("d,N", [(1, 300), (2, 513), (5, 256), (8, 1000), (10, 514), (16, 130), (17, 258), (40, 96),
                                 (64, 70), (100, 200), (128, 66)])
@pytest.mark.parametrize("kind", ["MarkovJumpHMC", "ContinuousTimeHMC", "ControlHMC", "HMC"])
def test_stream_kernel_matches_oracle(kind, d, N):
    """Every (warps per particle, dims per thread) plan; N chosen so the last tile is ragged."""
    dist_name = ("RoughWell", "TestGaussian", "Gaussian")[(d + len(kind)) % 3]
    hp = dict(epsilon=0.2, beta=0.3, num_leapfrog_steps=3)
    s, dist, o = _pair(kind, dist_name, d, N, 100 + d, hp)
    assert s._engine.fused and s._engine.kernel == "stream"
    X, Xo = s.sample(4, preserve_order=True), o.sample(4, preserve_order=True)
    assert X.shape == (d, N, 4)
    assert helpers.rel_err(X, Xo) < 1e-10
    assert _counters(s, dist) == [o.counters()[k] for k in KEYS]
    assert helpers.rel_err(s.state.V, o.V) < 1e-10


@pytest.mark.parametrize("N", [1, 31, 33, 255, 257, 1001])
def test_stream_kernel_odd_counts_take_the_cooperative_loader(N):
    """ld * 8 is not a multiple of 16 for odd N: no tensor map can describe the rows, the kernel loads them itself."""
    hp = dict(epsilon=0.25, beta=0.4, num_leapfrog_steps=2)
    s, dist, o = _pair("MarkovJumpHMC", "RoughWell", 6, N, 7 + N, hp)
    X, Xo = s.sample(3), o.sample(3)
    assert helpers.rel_err(X, Xo) < 1e-10
    assert _counters(s, dist) == [o.counters()[k] for k in KEYS]


@pytest.mark.parametrize("d", [2, 10, 100])
def test_tma_and_cooperative_loader_agree_bitwise(d):
    from mjhmc_b200 import _lib
    hp = dict(epsilon=0.3, beta=0.2, num_leapfrog_steps=4)
    lib = _lib.load()
    out = []
    for tma in (1, 0):
        lib.mjhmc_stream_set_tma(tma)
        try:
            s, dist, _ = _pair("MarkovJumpHMC", "RoughWell", d, 4096 + 64, 3, hp)
            out.append((s.sample(5), s.state.X.copy(), s.state.V.copy(), _counters(s, dist)))
        finally:
            lib.mjhmc_stream_set_tma(1)
    for a, b in zip(out[0][:3], out[1][:3]):
        np.testing.assert_array_equal(a, b)
    assert out[0][3] == out[1][3]


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("kind", ["MarkovJumpHMC", "ContinuousTimeHMC", "ControlHMC"])
@pytest.mark.parametrize("d", [2, 3, 8])
def test_stream_agrees_with_register_kernel(kind, d, dtype):
    """Same energy, same draws, two kernels.  Not bitwise: nvcc contracts x/s1^2 + c sin(..) into different FMAs in
    the two bodies (1 ulp), so fp64 is held to the 1e-10 of the oracle tests and the integer counters to equality;
    fp32 to the share of particles whose operator choices did not flip."""
    hp = dict(epsilon=0.3, beta=0.2, num_leapfrog_steps=5)
    res = []
    n_it = 6 if dtype == "float64" else 1          # fp32: one iteration, before 1-ulp differences are amplified
    for kernel in ("stream", "auto"):
        s, dist, _ = _pair(kind, "RoughWell", d, 3000, 11, hp, dtype=dtype, kernel=kernel)
        res.append((s.sample(n_it), s.state.V.copy(), _counters(s, dist), s.dwelling_times.copy()))
    if dtype == "float64":
        assert helpers.rel_err(res[0][0], res[1][0]) < 1e-10
        assert helpers.rel_err(res[0][1], res[1][1]) < 1e-10
        assert res[0][2] == res[1][2]
        fin = np.isfinite(res[1][3])
        assert np.array_equal(np.isfinite(res[0][3]), fin) and helpers.rel_err(res[0][3][fin], res[1][3][fin]) < 1e-9
    else:
        X, Xo = res[0][0], res[1][0]
        same = np.all(np.abs(X - Xo) <= 1e-3 * (1 + np.abs(Xo)), axis=0)
        assert same.mean() > 0.97 and helpers.rel_err32(X[:, same], Xo[:, same]) < 1e-4


def test_stream_one_launch_equals_many():
    hp = dict(epsilon=0.2, beta=0.3, num_leapfrog_steps=3)
    s1, d1, _ = _pair("MarkovJumpHMC", "Gaussian", 40, 1500, 5, hp)
    s2, d2, _ = _pair("MarkovJumpHMC", "Gaussian", 40, 1500, 5, hp)
    A = s1.sample(6, preserve_order=True)
    B = np.concatenate([s2.sample(1, preserve_order=True) for _ in range(6)], axis=2)
    np.testing.assert_array_equal(A, B)
    assert _counters(s1, d1) == _counters(s2, d2)
    assert s1._engine.launches == 1 and s2._engine.launches == 6


@pytest.mark.parametrize("name", [c for c in helpers.golden_inject_cases()
                                  if not c.endswith("backoff") and "multimodal" not in c])     # separable energies only
def test_stream_kernel_follows_the_reference_golden_trajectories(name):
    """Injected draws recorded from the unmodified reference (tests/golden/generate_golden.py), one launch."""
    g = helpers.load_inject(name)
    s, dist = helpers.product_from_golden(name, g, kernel="stream")
    n = g["X"].shape[0]
    X = s.sample(n)
    assert helpers.rel_err(X, np.concatenate(list(g["X"]), axis=1)) < 1e-10
    assert helpers.rel_err(s.state.V, g["V"][-1]) < 1e-10
    assert _counters(s, dist) == list(g["counters"][-1])


def test_stream_kernel_backoff_golden():
    """The reference's infinite-rate back-off (markov_jump_hmc.py:376-389) through the streaming kernel."""
    name = "MarkovJumpHMC_backoff"
    g = helpers.load_inject(name)
    s, dist = helpers.product_from_golden(name, g, kernel="stream")
    for it in range(g["X"].shape[0]):
        s.sampling_iteration()
        st = s.state
        assert helpers.rel_err(st.X, g["X"][it]) < 1e-10
        assert helpers.rel_err(st.V, g["V"][it]) < 1e-10
        assert _counters(s, dist) == list(g["counters"][it])
        assert np.array_equal(st.cache_active, g["cache"][it])
        s._host_state = None
    assert s.epsilon == float(g["final_epsilon"]) and s.num_leapfrog_steps == int(g["final_L"])


def test_reference_default_gaussian_100d_runs_on_the_streaming_kernel():
    """Gaussian(ndims=100) of the reference is diagonal (distributions.py:257-263): auto dispatch -> streaming kernel,
    searched hyper-parameters of search/MJHMC_log_gauss/params_2.json."""
    from mjhmc_b200.misc.distributions import Gaussian
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    from mjhmc_b200 import _lib
    rs = np.random.RandomState(0)
    d, N = 100, 2048
    dist = Gaussian(ndims=d, nbatch=N, log_conditioning=6)
    X0 = rs.randn(d, N) / np.sqrt(np.diag(dist.J))[:, None]
    V0 = rs.randn(d, N)
    helpers.pin_init(dist, X0)
    hp = dict(epsilon=1.4581446647644043, beta=0.009999999776482582, num_leapfrog_steps=25)
    s = MarkovJumpHMC(distribution=dist, V=V0, seed=2, resample=False, **hp)
    assert s._engine.desc.kind == _lib.DIST_DIAG_GAUSSIAN
    o = orc.OracleSampler("MarkovJumpHMC", orc.GaussianEnergy.log_conditioned(d, 6), X0, V=V0,
                          draws=orc.PhiloxDraws(2), resample=False, **hp)
    X, Xo = s.sample(3), o.sample(3)
    assert helpers.rel_err(X, Xo) < 1e-10
    assert _counters(s, dist) == [o.counters()[k] for k in KEYS]


@pytest.mark.parametrize("d", [20, 40, 100])
@pytest.mark.parametrize("kind", ["MarkovJumpHMC", "ControlHMC"])
def test_stream_kernel_fp32_with_several_threads_per_particle(kind, d):
    """fp32 states, G > 1: one iteration from the same state against the fp64 oracle (tolerance 1e-4 on the particles
    whose operator choice did not flip in single precision)."""
    dist_name = "Gaussian" if d != 40 else "RoughWell"
    hp = dict(epsilon=0.2, beta=0.3, num_leapfrog_steps=3)
    s, dist, o = _pair(kind, dist_name, d, 1500, 40 + d, hp, dtype="float32")
    X, Xo = s.sample(1), o.sample(1)
    same = np.all(np.abs(X - Xo) <= 1e-3 * (1 + np.abs(Xo)), axis=0)
    assert same.mean() > 0.97 and helpers.rel_err32(X[:, same], Xo[:, same]) < 1e-4


@pytest.mark.parametrize("d", [6, 24, 100])
def test_stream_kernel_backoff_with_several_threads_per_particle(d):
    """A non-finite rate inside a multi-iteration launch when the particle is spread over G warps: the failed
    particle must keep its state, the launch must be replayed and the batch-wide back-off of
    markov_jump_hmc.py:376-389 must give the oracle's trajectory and counters."""
    from mjhmc_b200.misc.distributions import TestGaussian
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    rs = np.random.RandomState(d)
    N = 70
    X0, V0 = rs.randn(d, N) * 0.3, rs.randn(d, N) * 0.1
    X0[:, 3] = 40.0                                     # H - H_L = 0.094 d x^2 > 709.78 at eps = 1, L = 1: exp() overflows
    dist = helpers.pin_init(TestGaussian(ndims=d, nbatch=N), X0)
    hp = dict(epsilon=1.0, beta=0.5, num_leapfrog_steps=1)
    s = MarkovJumpHMC(distribution=dist, V=V0, seed=21, resample=False, kernel="stream", **hp)
    o = orc.OracleSampler("MarkovJumpHMC", orc.TestGaussianEnergy(1.0), X0, V=V0, draws=orc.PhiloxDraws(21),
                          resample=False, **hp)
    X, Xo = s.sample(3), o.sample(3)
    assert helpers.rel_err(X, Xo) < 1e-10
    assert _counters(s, dist) == [o.counters()[k] for k in KEYS]
    assert s.epsilon == o.epsilon and s.num_leapfrog_steps == o.num_leapfrog_steps
    assert o.counters()["dEdX"] > N + 3 * 2 * N        # more than three plain iterations: the back-off ran
