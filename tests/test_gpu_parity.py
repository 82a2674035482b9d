"""GPU parity: the CUDA path (through the C ABI) against the golden reference trajectories
and against the oracle on the same injected draws.

Tolerances (north_star): trajectories 1e-10 relative in fp64, 1e-4 relative in fp32, per element
(helpers.rel_err: every value against its own magnitude, floored at one tenth of the array's RMS); integer counters
and operator choices exact.
"""
import numpy as np
import pytest

from oracle import mjhmc_oracle as orc
from tests import helpers

pytestmark = pytest.mark.gpu

TOL = {"float64": 1e-10, "float32": 1e-4}
CASES = [c for c in helpers.golden_inject_cases() if not c.endswith("backoff")]


def _counters(sampler, dist):
    return [sampler.l_count, sampler.f_count, sampler.fl_count, sampler.r_count, dist.E_count, dist.dEdX_count]


@pytest.mark.parametrize("name", CASES)
def test_golden_trajectory_per_iteration_fp64(name):
    """One launch per sampling_iteration(); state, energies, dwell times and counters after each."""
    g = helpers.load_inject(name)
    s, dist = helpers.product_from_golden(name, g)
    energy = helpers.energy_from_golden(g)
    assert s._engine.fused
    for it in range(g["X"].shape[0]):
        s.sampling_iteration()
        st = s.state
        assert helpers.rel_err(st.X, g["X"][it]) < TOL["float64"], (name, it)
        assert helpers.rel_err(st.V, g["V"][it]) < TOL["float64"], (name, it)
        # energies are sums of terms that cancel (sum of cosines), so their error does not follow their own magnitude:
        # positions that agree to TOL per element move E by at most sum_k |dE/dx_k| TOL |x_k| (first order), and that
        # -- the tolerance of the positions carried through the energy -- is what the energies are held to
        # (a fixed 1e-10 of the array's RMS sat at 0.9e-10 ... 1.0e-10 on the RoughWell case: rounding noise)
        e_tol = TOL["float64"] * (np.sum(np.abs(g["X"][it] * energy.dEdX(g["X"][it])), axis=0) + np.abs(g["EX"][it]).ravel())
        assert np.all(np.abs(st.EX[0] - g["EX"][it].ravel()) <= e_tol), (name, it)
        v_tol = TOL["float64"] * 2.0 * g["EV"][it].ravel()          # d(v^2/2) = v dv <= 2 EV TOL
        assert np.all(np.abs(st.EV[0] - g["EV"][it].ravel()) <= v_tol + 1e-300), (name, it)
        assert _counters(s, dist) == list(g["counters"][it]), (name, it)
        if name.startswith(("MarkovJumpHMC", "ContinuousTimeHMC")):
            fin = np.isfinite(g["dwell"][it])
            assert np.array_equal(np.isfinite(s.dwelling_times), fin)
            assert helpers.rel_err(s.dwelling_times[fin], g["dwell"][it][fin]) < 1e-9
        if name.startswith("MarkovJumpHMC"):
            assert np.array_equal(st.cache_active, g["cache"][it])
        s._host_state = None      # do not round-trip the state through the host


@pytest.mark.parametrize("name", CASES)
def test_golden_trajectory_single_launch_fp64(name):
    """sample(n): all iterations inside ONE persistent launch; every recorded sample column."""
    g = helpers.load_inject(name)
    s, dist = helpers.product_from_golden(name, g)
    n = g["X"].shape[0]
    X = s.sample(n)
    d, N = g["X0"].shape
    assert X.shape == (d, n * N) and X.dtype == np.float64
    expect = np.concatenate(list(g["X"]), axis=1)
    assert helpers.rel_err(X, expect) < TOL["float64"]
    assert helpers.rel_err(s.state.V, g["V"][-1]) < TOL["float64"]
    assert _counters(s, dist) == list(g["counters"][-1])
    assert s._engine.launches == 1


@pytest.mark.parametrize("name", CASES)
def test_single_iteration_from_oracle_state_fp32(name):
    """fp32: each iteration restarted from the oracle's fp64 state (no error accumulation)."""
    g = helpers.load_inject(name)
    o = helpers.oracle_from_golden(name, g)
    s, dist = helpers.product_from_golden(name, g, dtype="float32")
    mismatched = 0
    for it in range(g["X"].shape[0]):
        st = s.state
        st.X[:], st.V[:] = o.X, o.V
        st.cache_active[:], st.H_cache[:] = o.cache_active, o.H_cache
        s._attempt = o.attempt
        o.sampling_iteration()
        s.sampling_iteration()
        got = s.state
        # a float32 near-tie may pick another operator for a particle: count, do not compare those
        same = np.all(np.abs(got.X - o.X) <= 1e-3 * (1 + np.abs(o.X)), axis=0) & \
            np.all(np.abs(got.V - o.V) <= 1e-3 * (1 + np.abs(o.V)), axis=0)
        mismatched += int((~same).sum())
        assert helpers.rel_err32(got.X[:, same], o.X[:, same]) < TOL["float32"], (name, it)
        assert helpers.rel_err32(got.V[:, same], o.V[:, same]) < TOL["float32"], (name, it)
    assert mismatched <= 2, "too many operator-choice flips in float32: %d" % mismatched


def test_backoff_golden():
    """Infinite-rate back-off (markov_jump_hmc.py:376-389) against the reference trajectory."""
    name = "MarkovJumpHMC_backoff"
    g = helpers.load_inject(name)
    s, dist = helpers.product_from_golden(name, g)
    for it in range(g["X"].shape[0]):
        s.sampling_iteration()
        st = s.state
        assert helpers.rel_err(st.X, g["X"][it]) < 1e-10
        assert helpers.rel_err(st.V, g["V"][it]) < 1e-10
        assert _counters(s, dist) == list(g["counters"][it])
        assert s._attempt == int(g["attempts"][it])
        assert np.array_equal(st.cache_active, g["cache"][it])
        s._host_state = None
    assert s.epsilon == float(g["final_epsilon"]) and s.num_leapfrog_steps == int(g["final_L"])


def test_backoff_inside_multi_iteration_launch():
    """The same back-off when the failing iteration sits in the middle of one sample(n) launch."""
    name = "MarkovJumpHMC_backoff"
    g = helpers.load_inject(name)
    s, dist = helpers.product_from_golden(name, g)
    n = g["X"].shape[0]
    X = s.sample(n)
    assert helpers.rel_err(X, np.concatenate(list(g["X"]), axis=1)) < 1e-10
    assert _counters(s, dist) == list(g["counters"][-1])
    assert s._attempt == int(g["attempts"][-1])


def _backoff_cloud(N=4096, d=2):
    """TestGaussian at eps = 1, L = 1 with one particle (index 40) that runs into exp(dH) = inf at iterations 3
    and 4 (tests/test_gpu_multi.py uses the same cloud); the rest is a plain N(0, 1) cloud."""
    rs = np.random.RandomState(22)
    X0, V0 = rs.randn(d, 64), rs.randn(d, 64)
    X0[:, 40] = rs.randn(2) * 40
    V0[:, 40] = rs.randn(2) * 40
    rs2 = np.random.RandomState(23)
    return np.hstack((X0, rs2.randn(d, N - 64))), np.hstack((V0, rs2.randn(d, N - 64)))


@pytest.mark.parametrize("kernel", ["auto", "stream"])
def test_flf_energy_kept_after_an_F_move_is_dropped_when_epsilon_or_L_change(kernel):
    """The kernels keep H_L as the FLF energy of an F mover (FLF(F z) = F L z).  That shortcut only holds while
    (epsilon, L) stay the same: after the back-off iteration at (eps/2, 2L) (markov_jump_hmc.py:376-389) and when
    the caller edits the hyper-parameters between calls, the reference re-integrates (cache_active is False).
    4096 particles, so a stale energy would flip operator choices: samples, dwelling times and counters must
    follow the oracle through both events."""
    from mjhmc_b200.misc.distributions import TestGaussian
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    X0, V0 = _backoff_cloud()
    hp = dict(epsilon=1.0, beta=0.5, num_leapfrog_steps=1)
    dist = helpers.pin_init(TestGaussian(ndims=2, nbatch=X0.shape[1]), X0)
    s = MarkovJumpHMC(distribution=dist, V=V0, seed=5, resample=False, kernel=kernel, **hp)
    o = orc.OracleSampler("MarkovJumpHMC", orc.TestGaussianEnergy(1.0), X0, V=V0, draws=orc.PhiloxDraws(5),
                          resample=False, **hp)
    keys = ("l", "f", "fl", "r", "E", "dEdX")
    X, Xo = s.sample(7), o.sample(7)                      # back-offs at iterations 3 and 4, then two plain iterations
    assert o.attempt == 9 and s._attempt == 9
    assert helpers.rel_err(X, Xo) < 1e-10
    np.testing.assert_allclose(s.dwelling_times, o.dwelling_times, rtol=1e-9)
    assert _counters(s, dist) == [o.counters()[k] for k in keys]
    # the caller edits the hyper-parameters between two calls
    s.epsilon = o.epsilon = 0.6
    s.num_leapfrog_steps = o.num_leapfrog_steps = 3
    X, Xo = s.sample(4), o.sample(4)
    assert helpers.rel_err(X, Xo) < 1e-10
    np.testing.assert_allclose(s.dwelling_times, o.dwelling_times, rtol=1e-9)
    assert _counters(s, dist) == [o.counters()[k] for k in keys]


def test_failed_attempt_leaves_dwelling_times_untouched():
    """ContinuousTimeHMC raises on a non-finite rate (utils.py:41-48) before sampler.dwelling_times is assigned
    (markov_jump_hmc.py:266-274): the values of the last good iteration must survive the failed launch."""
    from mjhmc_b200.samplers.markov_jump_hmc import ContinuousTimeHMC
    from mjhmc_b200.misc.distributions import TestGaussian
    X0 = np.array([[.4, .1, .2, .3]])
    dist = helpers.pin_init(TestGaussian(ndims=1, nbatch=4), X0)
    c = ContinuousTimeHMC(distribution=dist, epsilon=1.0, beta=0.5, num_leapfrog_steps=1, V=np.zeros((1, 4)), seed=1,
                          resample=False)
    c.sample(3)
    before = c.dwelling_times.copy()
    assert np.all(before > 0)
    st = c.state
    st.X[0, 0], st.V[0, 0] = 100., 0.
    c.state = st
    with pytest.raises(ValueError, match="Infinite rate"):
        c.sample(2)
    np.testing.assert_array_equal(c.dwelling_times, before)


def test_continuous_time_raises_on_infinite_rate():
    from mjhmc_b200.samplers.markov_jump_hmc import ContinuousTimeHMC
    from mjhmc_b200.misc.distributions import TestGaussian
    dist = helpers.pin_init(TestGaussian(ndims=1, nbatch=4), np.array([[100., .1, .2, .3]]))
    c = ContinuousTimeHMC(distribution=dist, epsilon=1.0, beta=0.5, num_leapfrog_steps=1, V=np.zeros((1, 4)), seed=1)
    e0 = dist.E_count
    with pytest.raises(ValueError, match="Infinite rate"):
        c.sampling_iteration()
    assert dist.E_count - e0 == 4                      # the failed attempt's evaluations stay counted
    np.testing.assert_array_equal(c.state.X, [[100., .1, .2, .3]])


@pytest.mark.parametrize("kind", orc.KINDS)
@pytest.mark.parametrize("dist_name", ["RoughWell", "Funnel", "FunnelLiteral", "Gaussian10"])
def test_philox_mode_matches_oracle_with_numpy_philox(kind, dist_name):
    """PHILOX mode on the GPU == oracle fed by the numpy restatement of the same stream."""
    rs = np.random.RandomState(42)
    from mjhmc_b200.misc import distributions as D
    from mjhmc_b200.samplers import markov_jump_hmc as S
    N = 257
    if dist_name == "RoughWell":
        d, dist, energy = 2, D.RoughWell(2, N, scale1=5, scale2=4), orc.RoughWellEnergy(5, 4)
        X0 = rs.randn(d, N) * 3
    elif dist_name.startswith("Funnel"):
        lit = dist_name.endswith("Literal")
        d, dist, energy = 4, D.Funnel(scale=2.0, nbatch=N, ndims=4, literal_reference_energy=lit), orc.FunnelEnergy(2.0, lit)
        X0 = rs.randn(d, N) * 0.5
    else:
        d = 10
        dist, energy = D.Gaussian(ndims=d, nbatch=N, log_conditioning=1), orc.GaussianEnergy.log_conditioned(d, 1)
        X0 = rs.randn(d, N)
    V0 = rs.randn(d, N)
    helpers.pin_init(dist, X0)
    hp = dict(epsilon=0.2, beta=0.3, num_leapfrog_steps=3)
    n = 6
    if dist_name == "FunnelLiteral":
        # the graph as written is unbounded below: keep the horizon short enough to stay finite
        hp, n = dict(epsilon=0.02, beta=0.3, num_leapfrog_steps=2), 3
    extra = dict(resample=False) if kind in ("ContinuousTimeHMC", "MarkovJumpHMC") else {}
    seed, offset = 987654321, 1000
    s = getattr(S, kind)(distribution=dist, V=V0, seed=seed, particle_offset=offset, **hp, **extra)
    o = orc.OracleSampler(kind, energy, X0, V=V0, draws=orc.PhiloxDraws(seed, offset), resample=False, **hp)
    X = s.sample(n)
    Xo = o.sample(n)
    assert np.all(np.isfinite(Xo))
    assert helpers.rel_err(X, Xo) < 1e-9
    assert helpers.rel_err(s.state.V, o.V) < 1e-9
    c = o.counters()
    assert _counters(s, dist) == [c["l"], c["f"], c["fl"], c["r"], c["E"], c["dEdX"]]


def test_energy_gradient_kernels_match_oracle():
    from mjhmc_b200.misc import distributions as D
    rs = np.random.RandomState(3)
    W = rs.randn(7, 7) * 0.4
    nu = rs.rand(7) * 2 + 2.1
    J = rs.randn(6, 6)
    cases = [
        (D.RoughWell(3, 50, 7.0, 3.0), orc.RoughWellEnergy(7.0, 3.0), 3),
        (D.TestGaussian(4, 50, 1.7), orc.TestGaussianEnergy(1.7), 4),
        (D.Gaussian(5, 50, log_conditioning=3), orc.GaussianEnergy.log_conditioned(5, 3), 5),
        (D.Gaussian(6, 50, J=J), orc.GaussianEnergy(J), 6),
        (D.Gaussian(40, 50, log_conditioning=2), orc.GaussianEnergy.log_conditioned(40, 2), 40),
        (D.MultimodalGaussian(ndims=4, nbatch=50, separation=1), orc.MultimodalGaussianEnergy(1, 4), 4),
        (D.MultimodalGaussian(ndims=20, nbatch=50, separation=0.5), orc.MultimodalGaussianEnergy(0.5, 20), 20),
        (D.Funnel(scale=3.0, nbatch=50, ndims=6), orc.FunnelEnergy(3.0), 6),
        (D.Funnel(scale=1.5, nbatch=50, ndims=6, literal_reference_energy=True), orc.FunnelEnergy(1.5, True), 6),
        (D.ProductOfT(ndims=7, nbasis=7, nbatch=50, W=W, lognu=np.log(nu), b=rs.randn(7) * 0.1),
         orc.ProductOfTEnergy(W, nu.astype('float32'), (rs.randn(0) if False else None)), 7),
    ]
    for dist, energy, d in cases:
        X = rs.randn(d, 33)
        if isinstance(energy, orc.ProductOfTEnergy):
            energy = orc.ProductOfTEnergy(dist.weights, dist.nu, dist.bias)
        e0, g0 = dist.E_count, dist.dEdX_count
        E = dist.E(X)
        G = dist.dEdX(X)
        assert E.shape == (1, 33) and G.shape == (d, 33)
        assert (dist.E_count - e0, dist.dEdX_count - g0) == (33, 33)
        assert helpers.rel_err(E[0], energy.E(X)) < 1e-12, type(dist).__name__
        assert helpers.rel_err(G, energy.dEdX(X)) < 1e-12, type(dist).__name__


def _dense_case(dist_name, d, N, rs):
    from mjhmc_b200.misc import distributions as D
    if dist_name == "Gaussian":
        dist = D.Gaussian.rotated(ndims=d, nbatch=N, log_conditioning=2, seed=d)
        energy = orc.GaussianEnergy(dist.J)
        X0 = dist.Xinit.copy()
    else:
        W = (rs.randn(d, d) / np.sqrt(d)).astype(np.float32)
        nu = (rs.rand(d) * 2 + 2.1).astype(np.float32)
        b = (rs.randn(d) * 0.1).astype(np.float32)
        dist = D.ProductOfT(ndims=d, nbasis=d, nbatch=N, W=W, lognu=np.log(nu.astype(np.float64)), b=b)
        energy = orc.ProductOfTEnergy(dist.weights, dist.nu, dist.bias)
        X0 = rs.randn(d, N)
    return dist, energy, X0


@pytest.mark.parametrize("kind", orc.KINDS)
@pytest.mark.parametrize("dist_name,d", [("Gaussian", 5), ("Gaussian", 40), ("Gaussian", 100),
                                         ("ProductOfT", 6), ("ProductOfT", 36), ("ProductOfT", 100)])
def test_dense_kernel_matches_oracle(kind, dist_name, d):
    """K4 (DMMA): full-covariance Gaussian and ProductOfT, PHILOX mode, all sampler classes,
    particle counts that are not multiples of the warp / CTA tile."""
    from mjhmc_b200.samplers import markov_jump_hmc as S
    rs = np.random.RandomState(1000 + d)
    N = 77 if d < 100 else 150
    dist, energy, X0 = _dense_case(dist_name, d, N, rs)
    V0 = rs.randn(d, N)
    helpers.pin_init(dist, X0)
    hp = dict(epsilon=0.15, beta=0.3, num_leapfrog_steps=4)
    extra = dict(resample=False) if kind in ("ContinuousTimeHMC", "MarkovJumpHMC") else {}
    seed, offset = 4242 + d, 64
    s = getattr(S, kind)(distribution=dist, V=V0, seed=seed, particle_offset=offset, **hp, **extra)
    assert s._engine.fused
    o = orc.OracleSampler(kind, energy, X0, V=V0, draws=orc.PhiloxDraws(seed, offset), resample=False, **hp)
    n = 5
    X = s.sample(n)
    Xo = o.sample(n)
    assert helpers.rel_err(X, Xo) < 1e-9, (kind, dist_name, d)
    assert helpers.rel_err(s.state.V, o.V) < 1e-9
    c = o.counters()
    assert _counters(s, dist) == [c["l"], c["f"], c["fl"], c["r"], c["E"], c["dEdX"]]
    if kind in ("ContinuousTimeHMC", "MarkovJumpHMC"):
        fin = np.isfinite(o.dwelling_times)
        assert helpers.rel_err(s.dwelling_times[fin], o.dwelling_times[fin]) < 1e-8
    assert s._engine.launches == 1


@pytest.mark.parametrize("kind", orc.KINDS)
@pytest.mark.parametrize("dist_name,d,N", [("Gaussian", 12, 40), ("Gaussian", 40, 300), ("Gaussian", 100, 517),
                                           ("ProductOfT", 6, 40), ("ProductOfT", 36, 300), ("ProductOfT", 100, 517)])
def test_dense_float32_tcgen05(kind, dist_name, d, N):
    """fp32 states: the tcgen05 / TMEM / TMA kernel (bf16x3 operands, csrc/dense_tc.cu) against the fp64 oracle (1e-4),
    full-covariance Gaussian and ProductOfT, all sampler classes, several iterations in one launch."""
    from mjhmc_b200.samplers import markov_jump_hmc as S
    rs = np.random.RandomState(50 + d)
    dist, energy, X0 = _dense_case(dist_name, d, N, rs)
    V0 = rs.randn(d, N)
    helpers.pin_init(dist, X0)
    hp = dict(epsilon=0.1, beta=0.3, num_leapfrog_steps=3)
    extra = dict(resample=False) if kind in ("ContinuousTimeHMC", "MarkovJumpHMC") else {}
    s = getattr(S, kind)(distribution=dist, V=V0, seed=3, dtype="float32", particle_offset=7, **hp, **extra)
    assert s._engine.fused
    o = orc.OracleSampler(kind, energy, X0, V=V0, draws=orc.PhiloxDraws(3, 7), resample=False, **hp)
    n = 3
    X, Xo = s.sample(n), o.sample(n)
    assert X.shape == Xo.shape
    # a float32 near-tie may pick another operator for a particle: exclude those columns, but only a few
    same = np.all(np.abs(X - Xo) <= 1e-3 * (1 + np.abs(Xo)), axis=0)
    assert same.mean() > 0.97, same.mean()
    assert helpers.rel_err32(X[:, same], Xo[:, same]) < 1e-4
    c = o.counters()
    got = _counters(s, dist)
    assert got[4:] == [c["E"], c["dEdX"]] or same.mean() < 1.0
    assert s._engine.launches == 1


@pytest.mark.parametrize("dist_name", ["Gaussian", "ProductOfT"])
def test_dense_float32_single_iteration_from_oracle_state(dist_name):
    """fp32 tcgen05 kernel, MarkovJumpHMC, each iteration restarted from the oracle's fp64 state (no error accumulation,
    FLF cache hits and misses mixed inside a tile): trajectories to 1e-4, operator choices and counters exact up to
    float32 near-ties."""
    from mjhmc_b200.samplers import markov_jump_hmc as S
    d, N = 100, 700
    rs = np.random.RandomState(9)
    dist, energy, X0 = _dense_case(dist_name, d, N, rs)
    V0 = rs.randn(d, N)
    helpers.pin_init(dist, X0)
    hp = dict(epsilon=0.1, beta=0.3, num_leapfrog_steps=4)
    s = S.MarkovJumpHMC(distribution=dist, V=V0, seed=11, dtype="float32", resample=False, **hp)
    o = orc.OracleSampler("MarkovJumpHMC", energy, X0, V=V0, draws=orc.PhiloxDraws(11), resample=False, **hp)
    flips = 0
    for it in range(6):
        st = s.state
        st.X[:], st.V[:] = o.X, o.V
        st.cache_active[:], st.H_cache[:] = o.cache_active, o.H_cache
        s.state = st
        s._attempt = o.attempt
        o.sampling_iteration()
        _, _, ch = s._advance(1, want_choice=True)
        got = s.state
        same = ch[0].cpu().numpy() == o.last_choice
        flips += int((~same).sum())
        assert helpers.rel_err32(got.X[:, same], o.X[:, same]) < 1e-4, it
        assert helpers.rel_err32(got.V[:, same], o.V[:, same]) < 1e-4, it
        np.testing.assert_array_equal(got.cache_active[same], o.cache_active[same])
        s._host_state = None
    assert flips <= 6, flips


@pytest.mark.parametrize("kind", ["MarkovJumpHMC", "ContinuousTimeHMC", "ControlHMC"])
@pytest.mark.parametrize("dist_name,d,N", [("Gaussian", 100, 60000), ("ProductOfT", 36, 40004)])
def test_dense_float32_tma_boxes_equal_plain_loads_and_stores(kind, dist_name, d, N):
    """The tcgen05 kernel moves the state of a tile through a shared-memory stash with TMA tensor boxes (boxes that
    reach past a tile rewrite the unchanged state of the next tile's first particles, partial boxes at the end of a
    CTA's range leave with plain stores); without tensor maps the same stash is filled and emptied with plain loads and
    stores.  Several tiles per CTA, several iterations per launch, dwelling times and sample record: bit-identical."""
    from mjhmc_b200 import _lib
    from mjhmc_b200.samplers import markov_jump_hmc as S
    lib = _lib.load()
    hp = dict(epsilon=0.12, beta=0.4, num_leapfrog_steps=2)
    extra = dict(resample=False) if kind in ("ContinuousTimeHMC", "MarkovJumpHMC") else {}
    out = []
    _, _, X0 = _dense_case(dist_name, d, N, np.random.RandomState(9))
    V0 = np.random.RandomState(10).randn(d, N)
    for tma in (1, 0):
        dist, _, _ = _dense_case(dist_name, d, N, np.random.RandomState(9))
        helpers.pin_init(dist, X0)
        lib.mjhmc_stream_set_tma(tma)
        try:
            s = getattr(S, kind)(distribution=dist, V=V0, seed=21, dtype="float32", **hp, **extra)
            assert s._engine.fused
            X = s.sample(3)
            rec = [X, s.state.X.copy(), s.state.V.copy(), np.array(_counters(s, dist))]
            if extra:
                rec.append(np.asarray(s.dwelling_times).copy())
            X2 = s.sample(2)                      # a second launch starts from the state the first one stored
            rec.append(X2)
            out.append(rec)
        finally:
            lib.mjhmc_stream_set_tma(1)
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("dist_name,d", [("Gaussian", 40), ("ProductOfT", 36)])
def test_dense_float64_job_ranges_equal_eight_particle_groups(dist_name, d):
    """The DMMA kernel hands out ranges of 28-32 particles and packs the FLF trajectories of the particles that need
    them in front of the L trajectories (csrc/dense.cu); clouds too small to fill the GPU are cut into ranges of 8.
    One cloud of 60 000 particles (large ranges, range size adapting to the FLF fraction) against the same particles
    in three launches of 20 000 (ranges of 8 only): bit-identical samples, state, dwelling times and counters."""
    from mjhmc_b200.samplers import markov_jump_hmc as S
    N, n = 60000, 4
    rs = np.random.RandomState(12)
    _, _, X0 = _dense_case(dist_name, d, N, rs)
    V0 = rs.randn(d, N)
    hp = dict(epsilon=0.25, beta=0.5, num_leapfrog_steps=2)
    outs = []
    for bounds in ([(0, N)], [(0, 20000), (20000, 40000), (40000, N)]):
        parts, vs, dw, tot = [], [], [], np.zeros(6, dtype=np.int64)
        for lo, hi in bounds:
            dist, _, _ = _dense_case(dist_name, d, hi - lo, np.random.RandomState(12))
            helpers.pin_init(dist, X0[:, lo:hi])
            s = S.MarkovJumpHMC(distribution=dist, V=V0[:, lo:hi], seed=8, resample=False, particle_offset=lo, **hp)
            assert s._engine.fused
            parts.append(s.sample(n, preserve_order=True))
            vs.append(s.state.V.copy())
            dw.append(np.asarray(s.dwelling_times).copy())
            tot += np.array(_counters(s, dist))
        outs.append((np.concatenate(parts, axis=1), np.concatenate(vs, axis=1), np.concatenate(dw), tot))
    for a, b in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(a, b)
    assert outs[0][3][3] > 0 and outs[0][3][1] > 0          # R and F moves happened: FLF jobs were packed


@pytest.mark.parametrize("dist_name", ["Gaussian", "ProductOfT"])
def test_dense_float32_rows_do_not_depend_on_the_tile_packing(dist_name):
    """The tcgen05 kernel packs L jobs and the FLF jobs of uncached particles into 128-row tiles whose composition
    depends on the particle range of the CTA: the whole cloud and two unequal shards must give bit-identical samples
    (rows of an MMA are independent), with identical counters."""
    from mjhmc_b200.samplers import markov_jump_hmc as S
    d, N, n = 36, 1000, 6
    rs = np.random.RandomState(4)
    dist0, energy, X0 = _dense_case(dist_name, d, N, rs)
    V0 = rs.randn(d, N)
    hp = dict(epsilon=0.2, beta=0.3, num_leapfrog_steps=3)
    outs, cnts = [], []
    for bounds in ([(0, N)], [(0, 333), (333, N)]):
        parts, tot = [], np.zeros(6, dtype=np.int64)
        for lo, hi in bounds:
            dist, _, _ = _dense_case(dist_name, d, hi - lo, np.random.RandomState(4))
            helpers.pin_init(dist, X0[:, lo:hi])
            s = S.MarkovJumpHMC(distribution=dist, V=V0[:, lo:hi], seed=5, dtype="float32", resample=False,
                                particle_offset=lo, **hp)
            parts.append(s.sample(n, preserve_order=True))
            tot += np.array(_counters(s, dist))
        outs.append(np.concatenate(parts, axis=1))
        cnts.append(tot)
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(cnts[0], cnts[1])
