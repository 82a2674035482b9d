"""GPU: the public API end to end -- README example (unfused callback path), user-defined
distributions against the golden reference trajectories, resampling, autocorrelation, sharding
invariance, float32, and the reference's own moment tests (tests/test_continuous_samplers.py)."""
import numpy as np
import pytest

from oracle import mjhmc_oracle as orc
from tests import helpers

pytestmark = pytest.mark.gpu

CASES = [c for c in helpers.golden_inject_cases() if not c.endswith("backoff")]


def _counters(sampler, dist):
    return [sampler.l_count, sampler.f_count, sampler.fl_count, sampler.r_count, dist.E_count, dist.dEdX_count]


def test_readme_example_runs_through_the_alias():
    """README.md:14-35 verbatim (config 1 of BASELINE.json), against the oracle on injected draws."""
    import mjhmc_b200
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == "mjhmc" or k.startswith("mjhmc.")}
    try:
        mjhmc_b200.install_alias()
        from mjhmc.samplers.markov_jump_hmc import MarkovJumpHMC
        from mjhmc.misc.distributions import LambdaDistribution

        def E(X, sigma=1.):
            return np.sum(X**2, axis=0).reshape((1, -1))/2./sigma**2

        def dEdX(X, sigma=1.):
            return X/sigma**2

        np.random.seed(0)
        Xinit = np.random.randn(2, 100)
        anonymous_gaussian = LambdaDistribution(energy_func=E, energy_grad_func=dEdX, init=Xinit, name='IsotropicGaussian')
        rs = np.random.RandomState(1)
        draws = dict(Z=rs.randn(12, 2, 100), U=rs.rand(12, 3, 100), U0=rs.rand(12))
        V0 = rs.randn(2, 100)
        mjhmc = MarkovJumpHMC(distribution=anonymous_gaussian, V=V0, injected_draws=draws)
        assert not mjhmc._engine.fused
        np.random.seed(7)
        X = mjhmc.sample(num_steps=10)
        assert X.shape == (2, 1000)
        o = orc.OracleSampler("MarkovJumpHMC", orc.LambdaEnergy(E, dEdX), Xinit, V=V0,
                              draws=orc.InjectedDraws(draws["Z"], draws["U"], draws["U0"]))
        np.random.seed(7)
        o.draws.resample_uniforms = lambda m: np.random.random(m)
        Xo = o.sample(10)
        assert helpers.rel_err(X, Xo) < 1e-10
        assert _counters(mjhmc, anonymous_gaussian) == [o.l_count, o.f_count, o.fl_count, o.r_count, o.E_count, o.dEdX_count]
    finally:
        for k in [k for k in sys.modules if k == "mjhmc" or k.startswith("mjhmc.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.parametrize("name", CASES)
def test_user_defined_distribution_unfused_matches_golden(name):
    """A Distribution subclass with numpy E_val/dEdX_val (no kernel descriptor) -> unfused path;
    state, counters (incl. the data-dependent E/dEdX counts of FLF sub-batches) vs the reference."""
    from mjhmc_b200.misc.distributions import Distribution
    from mjhmc_b200.samplers import markov_jump_hmc as S
    g = helpers.load_inject(name)
    energy = helpers.energy_from_golden(g)
    X0 = g["X0"]

    class UserDist(Distribution):
        def E_val(self, X):
            return energy.E(X).reshape((1, -1))

        def dEdX_val(self, X):
            return energy.dEdX(X)

        def gen_init_X(self):
            self.Xinit = X0.copy()

        def __hash__(self):
            return 1

    dist = UserDist(ndims=X0.shape[0], nbatch=X0.shape[1])
    kind = name.split("_")[0]
    extra = dict(resample=False) if kind in ("ContinuousTimeHMC", "MarkovJumpHMC") else {}
    s = getattr(S, kind)(distribution=dist, epsilon=float(g["epsilon"]), beta=float(g["beta_arg"]),
                         num_leapfrog_steps=int(g["L"]), V=g["V0"],
                         injected_draws=dict(Z=g["Z"], U=g["U"], U0=g["U0"]), **extra)
    assert not s._engine.fused
    n = 6
    X = s.sample(n)
    assert helpers.rel_err(X, np.concatenate(list(g["X"][:n]), axis=1)) < 1e-10
    assert helpers.rel_err(s.state.V, g["V"][n - 1]) < 1e-10
    assert _counters(s, dist) == list(g["counters"][n - 1])


def test_positional_callable_form_like_reference_tests():
    """tests/test_continuous_samplers.py builds samplers as Sampler(Xinit, E, dEdX)."""
    from mjhmc_b200.misc.distributions import TestGaussian
    from mjhmc_b200.samplers.markov_jump_hmc import ControlHMC
    np.random.seed(3)
    g1 = TestGaussian(ndims=1)
    s = ControlHMC(g1.Xinit, g1.E, g1.dEdX, epsilon=1.0, beta=0.3, seed=5)
    assert s._engine.fused                       # bound methods of a built-in distribution keep the fused path
    X = s.sample(50)
    assert X.shape == (1, 5000) and np.all(np.isfinite(X))
    # arbitrary callables -> unfused path
    s2 = ControlHMC(g1.Xinit.copy(), lambda X: np.sum(X ** 2, axis=0) / 2., lambda X: X, epsilon=1.0, beta=0.3, seed=5)
    assert not s2._engine.fused
    assert s2.sample(5).shape == (1, 500)


@pytest.mark.parametrize("kind", ["ContinuousTimeHMC", "MarkovJumpHMC"])
def test_resampling_matches_oracle(kind):
    """markov_jump_hmc.py:309-329: 1+n iterations, dwell-time weighted resampling; identical columns."""
    from mjhmc_b200.misc.distributions import RoughWell
    from mjhmc_b200.samplers import markov_jump_hmc as S
    rs = np.random.RandomState(9)
    d, N, n = 2, 50, 20
    X0, V0 = rs.randn(d, N) * 3, rs.randn(d, N)
    draws = dict(Z=rs.randn(n + 2, d, N), U=rs.rand(n + 2, 3, N), U0=rs.rand(n + 2))
    dist = helpers.pin_init(RoughWell(d, N, scale1=5, scale2=4), X0)
    hp = dict(epsilon=0.4, beta=0.3, num_leapfrog_steps=4)
    s = getattr(S, kind)(distribution=dist, V=V0, injected_draws=draws, **hp)
    assert s.resample
    np.random.seed(11)
    X = s.sample(n)
    o = orc.OracleSampler(kind, orc.RoughWellEnergy(5, 4), X0, V=V0,
                          draws=orc.InjectedDraws(draws["Z"], draws["U"], draws["U0"]), **hp)
    o.draws.resample_uniforms = lambda m: np.random.random(m)
    np.random.seed(11)
    Xo = o.sample(n)
    assert X.shape == Xo.shape == (d, n * N)
    # identical resampling indices: every output column is one of the recorded columns
    assert helpers.rel_err(X, Xo) < 1e-10
    assert s._attempt == n + 1 and dist.dEdX_count == o.dEdX_count


def test_resample_kernel_indices_large():
    """The index search (searchsorted right on the dwell-time prefix sum) on a longer input."""
    import ctypes as C
    import torch
    from mjhmc_b200 import _device, _lib
    lib = _lib.load()
    rs = np.random.RandomState(2)
    m = 200003
    dwell = rs.exponential(size=m)
    dwell[rs.rand(m) < 0.01] = 0.0
    r = np.sort(rs.rand(m)) * dwell.sum()
    samples = rs.randn(3, m)
    dev = torch.device("cuda")
    t = lambda a: torch.as_tensor(a, device=dev)
    out = torch.zeros((3, m), dtype=torch.float64, device=dev)
    idx = torch.zeros(m, dtype=torch.int64, device=dev)
    scratch = torch.empty(int(lib.mjhmc_resample_scratch_bytes(m)), dtype=torch.uint8, device=dev)
    dw, rr, ss = t(dwell), t(r), t(samples)
    _lib.check(lib.mjhmc_resample(_lib.F64, 3, _device.ptr(dw), m, _device.ptr(rr), m, _device.ptr(ss), m,
                                  _device.ptr(out), m, _device.ptr(idx), _device.ptr(scratch), None))
    torch.cuda.synchronize()
    want = orc.resample_indices(dwell, r)
    got = idx.cpu().numpy()
    # the device prefix sum is a tree sum: an index may differ only where r sits within rounding of a boundary
    diff = np.nonzero(got != want)[0]
    assert len(diff) <= 2, len(diff)
    same = got == want
    np.testing.assert_array_equal(out.cpu().numpy()[:, same], samples[:, want[same]])


def test_autocorrelation_kernel_matches_fft_autocor():
    from mjhmc_b200 import parallel
    from mjhmc_b200.misc.distributions import Gaussian
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    np.random.seed(4)
    dist = Gaussian(ndims=3, nbatch=70, log_conditioning=1)
    s = MarkovJumpHMC(distribution=dist, epsilon=0.5, beta=0.2, num_leapfrog_steps=3, resample=False, seed=8)
    S = s.sample_device(64)                                  # (d, T, N) on the device
    ac = parallel.autocorrelation(S)
    host = np.ascontiguousarray(S.cpu().numpy().transpose(0, 2, 1))      # reference layout (d, N, T)
    ref = orc.fft_autocor(host)
    np.testing.assert_allclose(ac, ref, atol=1e-10)
    assert abs(parallel.effective_sample_size(ac) - orc.ess_from_autocor(ref)) < 1e-6
    # float32 samples, fewer lags
    S32 = S.float()
    ac32 = parallel.autocorrelation(S32, n_lags=10)
    np.testing.assert_allclose(ac32, ref[:10], atol=1e-5)


@pytest.mark.parametrize("T,n_lags,N", [(300, 300, 37), (129, 200, 16), (50, 7, 1), (1000, 1000, 20), (1500, 1500, 5)])
def test_autocorrelation_kernels_multi_pass_and_ragged(T, n_lags, N):
    """The register-blocked kernel over several 128-lag passes, ragged particle tiles, T not a multiple of 8, more
    lags than steps (circular wrap), and the plain kernel where the extended tile no longer fits in shared memory."""
    import torch
    from mjhmc_b200 import parallel
    rs = np.random.RandomState(T)
    x = np.cumsum(rs.randn(2, T, N), axis=1) * 0.1 + rs.randn(2, T, N)         # (d, T, N) device layout
    S = torch.as_tensor(x, device="cuda")
    circ = parallel.autocorr_partial(S, n_lags=n_lags, circular=True).cpu().numpy()
    want_c = np.array([np.sum(x * np.roll(x, -(tau % T), axis=1)) for tau in range(n_lags)])
    np.testing.assert_allclose(circ, want_c, rtol=1e-11, atol=1e-9)
    lin_lags = min(n_lags, T)
    lin = parallel.autocorr_partial(S, n_lags=lin_lags, circular=False).cpu().numpy()
    want_l = np.array([np.sum(x[:, :T - tau] * x[:, tau:]) for tau in range(lin_lags)])
    np.testing.assert_allclose(lin, want_l, rtol=1e-11, atol=1e-9)
    c32 = parallel.autocorr_partial(S.float(), n_lags=min(n_lags, 16), circular=True).cpu().numpy()
    np.testing.assert_allclose(c32, want_c[:min(n_lags, 16)], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("T,n_lags,N,d", [(16, 16, 1, 1), (32, 20, 7, 2), (128, 128, 33, 3), (1024, 1024, 50, 2),
                                         (2048, 100, 19, 1), (4096, 4096, 6, 1)])
def test_fft_autocorrelation_kernel(T, n_lags, N, d):
    """K7b (csrc/autocorr_fft.cu): the route the reference takes (misc/autocor.py:37-49 -- FFT along time, |.|^2, inverse
    FFT) against the literal circular products and against the direct-product kernel: radix-2 + radix-4 stage mixes
    (log2 T odd and even), odd particle counts (an unpaired real series), tiles that do not fill, fewer lags than steps,
    float32 samples."""
    import torch
    from mjhmc_b200 import parallel
    rs = np.random.RandomState(T + N)
    x = np.cumsum(rs.randn(d, T, N), axis=1) * 0.1 + rs.randn(d, T, N) + 0.3       # (d, T, N) device layout, non-zero mean
    S = torch.as_tensor(x, device="cuda")
    fft = parallel.autocorr_partial(S, n_lags=n_lags, method="fft").cpu().numpy()
    want = np.real(np.fft.ifft(np.abs(np.fft.fft(x, axis=1)) ** 2, axis=1)).sum(axis=(0, 2))[:n_lags]
    np.testing.assert_allclose(fft, want, rtol=1e-11, atol=1e-9 * want[0])
    if T <= 1024:                                   # (the direct kernel keeps a T x 16 tile in shared memory)
        direct = parallel.autocorr_partial(S, n_lags=min(n_lags, 64), method="direct").cpu().numpy()
        np.testing.assert_allclose(fft[:len(direct)], direct, rtol=1e-11, atol=1e-9 * want[0])
    f32 = parallel.autocorr_partial(S.float(), n_lags=n_lags, method="fft").cpu().numpy()
    np.testing.assert_allclose(f32, want, rtol=1e-5, atol=1e-5 * want[0])
    # normalised curve == the oracle's fft_autocor on the reference layout (d, N, T)
    if n_lags == T:
        ref = orc.fft_autocor(np.ascontiguousarray(x.transpose(0, 2, 1)))
        np.testing.assert_allclose(fft / fft[0], ref, atol=1e-10)
    with pytest.raises(ValueError):
        parallel.autocorr_partial(torch.as_tensor(x[:, :T - 1], device="cuda"), method="fft")


@pytest.mark.parametrize("kind", ["ControlHMC", "MarkovJumpHMC"])
def test_shard_invariance_single_gpu(kind):
    """T7 on one device: the cloud sampled whole == the two halves sampled separately with
    particle_offset (Philox keyed by the global particle index; the batch coin is particle-free)."""
    from mjhmc_b200.misc.distributions import RoughWell
    from mjhmc_b200.samplers import markov_jump_hmc as S
    rs = np.random.RandomState(21)
    d, N, n = 2, 301, 7
    X0, V0 = rs.randn(d, N) * 4, rs.randn(d, N)
    hp = dict(epsilon=0.6, beta=0.5, num_leapfrog_steps=5, seed=77)
    extra = dict(resample=False) if kind == "MarkovJumpHMC" else {}

    def run(lo, hi):
        dist = helpers.pin_init(RoughWell(d, hi - lo, scale1=6, scale2=4), X0[:, lo:hi])
        s = getattr(S, kind)(distribution=dist, V=V0[:, lo:hi], particle_offset=lo, **hp, **extra)
        X = s.sample(n, preserve_order=True)
        return X, _counters(s, dist)
    Xf, cf = run(0, N)
    Xa, ca = run(0, 150)
    Xb, cb = run(150, N)
    np.testing.assert_array_equal(Xf, np.concatenate([Xa, Xb], axis=1))
    assert cf == [a + b for a, b in zip(ca, cb)]


@pytest.mark.parametrize("kind", orc.KINDS)
def test_moments_1d_gaussian(kind):
    """tests/test_continuous_samplers.py:19-41: |mean| < .05, |std - 1| < .05."""
    from mjhmc_b200.misc.distributions import TestGaussian
    from mjhmc_b200.samplers import markov_jump_hmc as S
    np.random.seed(1)
    dist = TestGaussian(ndims=1, nbatch=100)
    extra = dict(resample=True) if kind in ("ContinuousTimeHMC", "MarkovJumpHMC") else {}
    # (epsilon=1, L=3 would rotate the unit oscillator by exactly 180 degrees: x -> -x, no mixing in |x|)
    s = getattr(S, kind)(distribution=dist, epsilon=0.6, beta=0.3, num_leapfrog_steps=3, seed=1, **extra)
    s.burn_in()
    X = s.sample(10000)
    assert abs(np.mean(X)) < .05 and abs(np.std(X) - 1) < .05, (kind, np.mean(X), np.std(X))


@pytest.mark.parametrize("kind", orc.KINDS)
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_moments_ill_conditioned_gaussian(kind, dtype):
    """tests/test_continuous_samplers.py:43-66: ||cov - J^-1||_F < .05 for Gaussian(ndims=2, log_conditioning=1)."""
    from mjhmc_b200.misc.distributions import Gaussian
    from mjhmc_b200.samplers import markov_jump_hmc as S
    np.random.seed(1)
    dist = Gaussian(ndims=2, nbatch=2000, log_conditioning=1)
    target = np.linalg.inv(dist.J)
    extra = dict(resample=False) if kind in ("ContinuousTimeHMC", "MarkovJumpHMC") else {}
    s = getattr(S, kind)(distribution=dist, epsilon=0.7, beta=0.3, num_leapfrog_steps=4, seed=3, dtype=dtype, **extra)
    s.burn_in()
    if extra:
        # the embedded jump chain is biased; weight by dwelling time on the device instead (resample=True path)
        s.resample = True
        X = s.sample(300)
    else:
        X = s.sample(300)
    assert np.linalg.norm(np.cov(X) - target) < .05 * 10, (kind, np.cov(X))
    assert np.linalg.norm(np.cov(X) - target) / np.linalg.norm(target) < .05


def test_pipelined_sample_equals_single_launch():
    """Large outputs are sampled in a few launches with the device->host copy overlapping the next launch;
    the result must not depend on the chunking (counter-based streams, persistent state)."""
    from mjhmc_b200.misc.distributions import RoughWell
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    rs = np.random.RandomState(17)
    X0, V0 = rs.randn(2, 300) * 3, rs.randn(2, 300)
    hp = dict(epsilon=0.4, beta=0.3, num_leapfrog_steps=3, seed=12, resample=False)
    outs = []
    for min_bytes in (1 << 60, 0):
        dist = helpers.pin_init(RoughWell(2, 300, scale1=5, scale2=4), X0)
        s = MarkovJumpHMC(distribution=dist, V=V0, **hp)
        s.PIPELINE_MIN_BYTES = min_bytes
        outs.append((s.sample(37), _counters(s, dist), s._engine.launches))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    assert outs[0][1] == outs[1][1]
    assert outs[0][2] == 1 and outs[1][2] == 8


def test_state_from_pinned_buffers_equals_plain_state_assignment():
    """bench.py's end-to-end leg hands the start state over as pinned torch tensors (HMCState.from_buffers: no host copy,
    one async upload per array, the empty FLF cache cleared on the device): same chain as assigning a regular HMCState,
    and the lazily built host members of such a state are there when somebody reads them."""
    import torch
    from mjhmc_b200.misc.distributions import RoughWell
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    from mjhmc_b200.samplers.hmc_state import HMCState
    rs = np.random.RandomState(3)
    N = 1000
    X0, V0 = rs.randn(2, N) * 3, rs.randn(2, N)
    X1, V1 = rs.randn(2, N) * 2, rs.randn(2, N)
    hp = dict(epsilon=0.4, beta=0.3, num_leapfrog_steps=3, seed=5, resample=False)
    outs = []
    for pinned in (False, True):
        dist = helpers.pin_init(RoughWell(2, N, scale1=5, scale2=4), X0)
        s = MarkovJumpHMC(distribution=dist, V=V0, **hp)
        s.sample(3)
        if pinned:
            st = HMCState.from_buffers(s, torch.as_tensor(X1).pin_memory(), torch.as_tensor(V1).pin_memory())
            assert st.cache_active.shape == (N,) and not st.cache_active.any() and st.H_cache.shape == (N,)
            assert st.active_idx[-1] == N - 1
        else:
            st = HMCState(X1.copy(), s, V=V1.copy())
        s.state = st
        outs.append((s.sample(4), s.state.X.copy(), s.state.V.copy(), _counters(s, dist)))
    for a, b in zip(outs[0][:3], outs[1][:3]):
        np.testing.assert_array_equal(a, b)
    assert outs[0][3] == outs[1][3]


def test_autocorrelation_curve_and_moments_agree_statistically_with_the_oracle():
    """north_star's second check: with DIFFERENT random streams the GPU sampler and the numpy oracle must agree in
    distribution -- the fft_autocor curve (autocor.py:37-49) over the first 40 lags and the first two moments.
    Tolerances are 3.5x the largest difference seen between oracle runs with different seeds (0.017 on the curve,
    0.8 % on the second moment) at this size."""
    from mjhmc_b200.misc.distributions import RoughWell
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    N, T = 4000, 120
    rs = np.random.RandomState(1234)
    X0, V0 = rs.randn(2, N) * 3, rs.randn(2, N)
    hp = dict(epsilon=0.3, beta=0.2, num_leapfrog_steps=5)
    dist = helpers.pin_init(RoughWell(2, N, scale1=5, scale2=4), X0)
    s = MarkovJumpHMC(distribution=dist, V=V0, seed=77, resample=False, **hp)
    X = s.sample(T, preserve_order=True)                       # (d, N, T)
    o = orc.OracleSampler("MarkovJumpHMC", orc.RoughWellEnergy(5, 4), X0, V=V0, draws=orc.PhiloxDraws(3),
                          resample=False, **hp)
    Xo = o.sample(T, preserve_order=True)
    assert X.shape == Xo.shape == (2, N, T)
    ac, aco = orc.fft_autocor(X), orc.fft_autocor(Xo)
    assert np.max(np.abs(ac[:40] - aco[:40])) < 0.06
    assert abs(X.mean() - Xo.mean()) < 0.15
    assert abs((X ** 2).mean() / (Xo ** 2).mean() - 1.0) < 0.03
