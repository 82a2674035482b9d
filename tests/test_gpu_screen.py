"""The single-precision screening of the holding-time race (csrc/common.cuh: decide_mj_screened / decide_ct_screened)
must never change a result: the operator choice is the one the literal fp64 code makes (misc/utils.py:15-49,
markov_jump_hmc.py:261-275, :366-396) and every stored holding time is the literal expression.

Two kinds of evidence:
  * the transition kernel called through the C ABI on crafted (energy difference, uniform) tuples -- near-ties placed
    at every distance from 1e-12 to 1e-2 around the decision boundaries, extreme energy differences, zero rates --
    against the numpy restatement of the race, and against the same call with MJHMC_RNG_FLAG_LITERAL_RACE;
  * whole sampler runs (every kernel family, Philox streams, the benchmark's hyper-parameters) with and without the
    screen: positions, momenta, counters and holding times bit-identical.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


# ---------------------------------------------------------------------------------------------------------------
# numpy restatement of the race
# ---------------------------------------------------------------------------------------------------------------
def _exp_draw(rate, u):
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        t = (1.0 / rate) * (-np.log(1.0 - u))
    return np.where(rate == 0.0, np.inf, t)


def _first_min(a, b, c):
    """utils.py:15-35 min_idx over three rows: the first minimum wins."""
    choice = np.zeros(a.shape, dtype=np.int64)
    dwell = a.copy()
    m = b < dwell
    choice[m] = 1
    dwell[m] = b[m]
    m = c < dwell
    choice[m] = 2
    dwell[m] = c[m]
    return choice, dwell


def race_ct(p_r, u0, u1, u2, e_fl):
    with np.errstate(over="ignore"):
        rfl = np.exp(e_fl) ** .5
    return _first_min(_exp_draw(np.ones_like(u1), u1), _exp_draw(rfl, u0), _exp_draw(np.full_like(u2, p_r), u2))


def race_mj(p_r, u0, u1, u2, e_l, e_flf):
    with np.errstate(over="ignore"):
        rl, rflf = np.exp(e_l) ** .5, np.exp(e_flf) ** .5
    rf = rflf - np.minimum(rl, rflf)
    return _first_min(_exp_draw(rl, u0), _exp_draw(rf, u1), _exp_draw(np.full_like(u2, p_r), u2))


# ---------------------------------------------------------------------------------------------------------------
# the transition kernel through the C ABI
# ---------------------------------------------------------------------------------------------------------------
def _transition(sampler_code, p_r, U, e_l, e_flf=None, literal=False):
    """One mjhmc_transition call on n one-dimensional particles whose energies are set so that H - H_L = e_l and
    H - H_FLF = e_flf.  Returns (choice, dwell)."""
    from mjhmc_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    n = e_l.shape[0]
    f64 = dict(dtype=torch.float64, device=dev)
    z = lambda *s: torch.zeros(*s, **f64)
    EX = torch.as_tensor(e_l, **f64).clone()
    cur = _lib.FullState(z(1, n).data_ptr(), z(1, n).data_ptr(), z(1, n).data_ptr(), EX.data_ptr(), z(n).data_ptr())
    keep = [z(1, n), z(1, n), z(1, n), z(n), z(n)]
    prop = _lib.FullState(*[t.data_ptr() for t in keep])
    H_flf = None
    Hc, ca = z(n), torch.zeros(n, dtype=torch.uint8, device=dev)
    if e_flf is not None:
        H_flf = torch.as_tensor(e_l - e_flf, **f64)           # the kernel forms H - H_flf = e_l - (e_l - e_flf)
    hp = _lib.HP()
    hp.sampler, hp.num_leapfrog_steps, hp.epsilon, hp.beta, hp.p_flip, hp.p_r = sampler_code, 1, 0.1, 0.5, 0.5, p_r
    Ud = torch.as_tensor(np.ascontiguousarray(U.reshape(1, 3, n)), **f64)
    Zd = z(1, 1, n)
    rng = _lib.RNG()
    rng.mode, rng.flags, rng.seed, rng.attempt0, rng.particle0 = _lib.RNG_INJECT, int(literal), 0, 0, 0
    rng.Z, rng.U, rng.U0, rng.inj_ld = Zd.data_ptr(), Ud.data_ptr(), None, n
    tmpl = np.zeros((_lib.COUNTER_ROWS, _lib.N_COUNTERS), dtype=np.int64)
    tmpl[:_lib.COUNTER_STRIPES, _lib.CNT_FAIL] = _lib.INT64_MAX
    counters = torch.as_tensor(tmpl, device=dev)
    dwell = z(n)
    choice = torch.zeros(n, dtype=torch.uint8, device=dev)
    o = _lib.Outputs()
    o.dwell, o.choice, o.counters = dwell.data_ptr(), choice.data_ptr(), counters.data_ptr()
    rc = lib.mjhmc_transition(_lib.F64, 1, C.byref(hp), C.byref(rng), n, n, C.byref(cur), C.byref(prop),
                              C.c_void_p(H_flf.data_ptr() if H_flf is not None else 0),
                              C.c_void_p(Hc.data_ptr()), C.c_void_p(ca.data_ptr()), C.byref(o), None)
    _lib.check(rc, "transition")
    torch.cuda.synchronize()
    out = (C.c_int64 * _lib.N_COUNTERS)()
    _lib.check(lib.mjhmc_counters_read(C.c_void_p(counters.data_ptr()), out, None), "counters_read")
    assert out[_lib.CNT_FAIL] == _lib.INT64_MAX, "a crafted case reported a non-finite rate"
    return choice.cpu().numpy().astype(np.int64), dwell.cpu().numpy()


def _near_tie_cases(rng, n, mj):
    """Uniforms and energy differences whose second holding time sits at a relative distance delta from the first,
    delta log-uniform in [1e-12, 1e-2] with both signs -- around and far inside the band the screen cannot decide."""
    u0, u1, u2 = rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    small = rng.uniform(0, 1, n) < 0.2                      # small uniforms: -log(1 - u) ~ u, the ill-conditioned end
    u0[small] = 10 ** rng.uniform(-12, -2, small.sum())
    small = rng.uniform(0, 1, n) < 0.2
    u1[small] = 10 ** rng.uniform(-12, -2, small.sum())
    delta = 10 ** rng.uniform(-12, -2, n) * rng.choice([-1.0, 1.0], n)
    w0, w1 = -np.log1p(-u0), -np.log1p(-u1)
    if not mj:
        # t_fl = w0 / r = t_f (1 + delta) = w1 (1 + delta)  =>  r = w0 / (w1 (1 + delta)),  e = 2 log r
        e = 2 * np.log(w0 / (w1 * (1 + delta)))
        return u0, u1, u2, np.clip(e, -60, 60), None
    # MarkovJumpHMC: t_l = w0 / r_l against t_f = w1 / (r_flf - r_l)
    e_l = rng.uniform(-20, 20, n)
    rl = np.exp(e_l / 2)
    rf = rl * w1 / (w0 * (1 + delta))                       # makes t_f = t_l (1 + delta)
    e_flf = 2 * np.log(rl + rf)
    return u0, u1, u2, e_l, np.clip(e_flf, -60, 60)


def _wild_cases(rng, n, mj):
    """Energy differences across the whole range the screen gates on (|e| around 64, hundreds, zero rates after
    underflow) and uniforms at the ends of [0, 1)."""
    u = rng.uniform(0, 1, (3, n))
    edge = rng.integers(0, 6, (3, n))
    u[edge == 0] = 0.0
    u[edge == 1] = 1.0 - 2.0 ** -53
    u[edge == 2] = 2.0 ** -53
    e_l = rng.choice([-800.0, -745.2, -700.0, -64.0001, -64.0, -63.9999, -1.0, 0.0, 1e-9, 1.0, 63.9999, 64.0, 64.0001,
                      300.0, 700.0], n) + rng.choice([0.0, 1e-7, 0.3], n)
    e_flf = None
    if mj:
        e_flf = e_l + rng.choice([-1e-9, 0.0, 1e-9, 1e-3, -1e-3, 5.0, -5.0, 100.0, -100.0], n)
        e_flf = np.minimum(e_flf, 705.0)
    return u[0], u[1], u[2], np.minimum(e_l, 705.0), e_flf


@pytest.mark.parametrize("maker", [_near_tie_cases, _wild_cases], ids=["near_ties", "wild"])
@pytest.mark.parametrize("p_r", [0.0, 0.2, 37.5])
@pytest.mark.parametrize("mj", [False, True], ids=["ContinuousTimeHMC", "MarkovJumpHMC"])
def test_screened_race_equals_the_literal_race(mj, p_r, maker):
    from mjhmc_b200 import _lib
    rng = np.random.default_rng(11 + int(mj) + int(p_r * 10))
    n = 1 << 20
    u0, u1, u2, e_l, e_flf = maker(rng, n, mj)
    if p_r == 0.0:
        u2 = np.zeros(n)                                      # the kernels do not draw u2 when the rate is zero
    U = np.stack([u0, u1, u2])
    code = _lib.SAMPLER_MARKOV_JUMP if mj else _lib.SAMPLER_CONTINUOUS_TIME
    c_s, d_s = _transition(code, p_r, U, e_l, e_flf, literal=False)
    c_l, d_l = _transition(code, p_r, U, e_l, e_flf, literal=True)
    np.testing.assert_array_equal(c_s, c_l)
    np.testing.assert_array_equal(d_s.view(np.int64), d_l.view(np.int64))          # bit-identical holding times
    # and both against numpy.  glibc's exp / log differ from CUDA's in the last ulp, so a tie closer than that noise may
    # fall either way: every disagreement must be such a tie (for MarkovJumpHMC the noise is amplified by the
    # cancellation in r_f = r_flf - r_l, markov_jump_hmc.py:368), and there must be few of them
    if mj:
        e_flf_k = e_l - (e_l - e_flf)
        c_ref, d_ref = race_mj(p_r, u0, u1, u2, e_l, e_flf_k)
        with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
            rl, rflf = np.exp(e_l) ** .5, np.exp(e_flf_k) ** .5
            amp = np.where(rflf > rl, rflf / (rflf - rl), 1.0)
        amp = np.where(np.isfinite(amp), amp, 1.0)
    else:
        c_ref, d_ref = race_ct(p_r, u0, u1, u2, e_l)
        amp = np.ones(n)
    bad = c_s != c_ref
    if bad.any():
        if mj:
            with np.errstate(over="ignore"):
                rl, rflf = np.exp(e_l) ** .5, np.exp(e_flf_k) ** .5
            t = np.stack([_exp_draw(rl, u0), _exp_draw(rflf - np.minimum(rl, rflf), u1), _exp_draw(np.full(n, p_r), u2)])
        else:
            with np.errstate(over="ignore"):
                rfl = np.exp(e_l) ** .5
            t = np.stack([_exp_draw(np.ones(n), u1), _exp_draw(rfl, u0), _exp_draw(np.full(n, p_r), u2)])
        ts = np.sort(t[:, bad], axis=0)
        with np.errstate(invalid="ignore"):
            gap = (ts[1] - ts[0]) / ts[0]
        assert np.all((gap <= 1e-14 * amp[bad]) | ~np.isfinite(gap)), "a disagreement with numpy that is not a tie"
        assert int(bad.sum()) <= n // 1000
    ok = ~bad & np.isfinite(d_ref)
    np.testing.assert_array_equal(np.isfinite(d_s[~bad]), np.isfinite(d_ref[~bad]))
    assert np.all(np.abs(d_s[ok] - d_ref[ok]) <= 4e-15 * amp[ok] * np.abs(d_ref[ok]))


# ---------------------------------------------------------------------------------------------------------------
# whole runs, every kernel family
# ---------------------------------------------------------------------------------------------------------------
def _run(make, literal, n_iter, resample):
    np.random.seed(20161017)            # the initial momenta and the resampling uniforms come from np.random
    sampler = make(literal, resample)
    S = sampler.sample(n_iter)
    st = sampler.state
    counts = (sampler.l_count, sampler.f_count, sampler.fl_count, sampler.r_count, sampler.distribution.E_count,
              sampler.distribution.dEdX_count)
    dw = np.array(sampler.dwelling_times, dtype=np.float64) if hasattr(sampler, "dwelling_times") else None
    return S, np.array(st.X), np.array(st.V), counts, dw


def _families():
    from mjhmc_b200.misc import distributions as D
    from mjhmc_b200.samplers import markov_jump_hmc as S
    from tests.helpers import pin_init

    def rough_well(kind, n):            # the headline hyper-parameters: unstable leapfrog, |H - H'| up to thousands
        def make(literal, resample):
            rng = np.random.default_rng(5)
            dist = pin_init(D.RoughWell(ndims=2, nbatch=n), 100 * rng.standard_normal((2, n)))
            return getattr(S, kind)(distribution=dist, epsilon=3.0, beta=0.0123144, num_leapfrog_steps=25, seed=7,
                                    resample=resample, literal_race=literal)
        return make

    def funnel(kind, n):                # BASELINE config 5
        def make(literal, resample):
            rng = np.random.default_rng(6)
            x0 = 3.0 * rng.standard_normal((1, n))
            X0 = np.concatenate((x0, np.exp(x0 / 2.) * rng.standard_normal((9, n))))
            dist = pin_init(D.Funnel(scale=3.0, ndims=10, nbatch=n), X0)
            return getattr(S, kind)(distribution=dist, epsilon=0.1, beta=0.5, num_leapfrog_steps=10, seed=8,
                                    resample=resample, literal_race=literal)
        return make

    def diag_gauss(kind, n, kernel):    # streaming kernel, several threads per particle
        def make(literal, resample):
            rng = np.random.default_rng(7)
            dist = pin_init(D.Gaussian(ndims=100, nbatch=n, log_conditioning=2), rng.standard_normal((100, n)))
            return getattr(S, kind)(distribution=dist, epsilon=0.3, beta=0.2, num_leapfrog_steps=5, seed=9,
                                    resample=resample, literal_race=literal, kernel=kernel)
        return make

    def full_gauss(kind, n, dtype):     # DMMA (fp64) and tcgen05 (fp32) dense kernels
        def make(literal, resample):
            rng = np.random.default_rng(8)
            A = rng.standard_normal((24, 24))
            J = A @ A.T / 24 + np.eye(24)
            dist = pin_init(D.Gaussian(ndims=24, nbatch=n, J=J), rng.standard_normal((24, n)))
            return getattr(S, kind)(distribution=dist, epsilon=0.25, beta=0.3, num_leapfrog_steps=4, seed=10,
                                    resample=resample, literal_race=literal, dtype=dtype)
        return make

    def callback(kind, n):              # unfused pieces around a Python energy
        def make(literal, resample):
            rng = np.random.default_rng(9)
            dist = pin_init(D.LambdaDistribution(energy_func=lambda X: np.sum(X ** 4, axis=0).reshape(1, -1) / 4.,
                                                 energy_grad_func=lambda X: X ** 3, init=rng.standard_normal((3, n)),
                                                 name="quartic"), rng.standard_normal((3, n)))
            return getattr(S, kind)(distribution=dist, epsilon=0.2, beta=0.4, num_leapfrog_steps=3, seed=11,
                                    resample=resample, literal_race=literal)
        return make

    fam = []
    for kind in ("ContinuousTimeHMC", "MarkovJumpHMC"):
        fam += [("fused-roughwell2d-%s" % kind, rough_well(kind, 200_000), 24),
                ("fused-funnel10d-%s" % kind, funnel(kind, 100_000), 24),
                ("stream-gauss100d-%s" % kind, diag_gauss(kind, 30_000, "stream"), 12),
                ("dense-f64-%s" % kind, full_gauss(kind, 20_000, "float64"), 12),
                ("dense-tc-f32-%s" % kind, full_gauss(kind, 20_000, "float32"), 12),
                ("unfused-%s" % kind, callback(kind, 5_000), 6)]
    return fam


def _family_ids():
    return [f[0] for f in _families()]


@pytest.mark.parametrize("resample", [False, True], ids=["last_dwell_only", "dwell_recorded"])
@pytest.mark.parametrize("idx", range(12), ids=lambda i: "family%d" % i)
def test_runs_are_bit_identical_with_and_without_the_screen(idx, resample):
    name, make, n_iter = _families()[idx]
    a = _run(make, False, n_iter, resample)
    b = _run(make, True, n_iter, resample)
    assert a[3] == b[3], name                                               # counters
    for x, y in zip(a[:3], b[:3]):
        np.testing.assert_array_equal(np.asarray(x).view(np.int64), np.asarray(y).view(np.int64), err_msg=name)
    if a[4] is not None:
        np.testing.assert_array_equal(a[4].view(np.int64), b[4].view(np.int64), err_msg=name)


# ---------------------------------------------------------------------------------------------------------------
# state in shared memory (fused_stash_kernel, ndims >= 6) against the register-resident kernel
# ---------------------------------------------------------------------------------------------------------------
def _stash_case(dist_name, d, dtype, kind, register_state, n=40_000, n_iter=10):
    from mjhmc_b200.misc import distributions as D
    from mjhmc_b200.samplers import markov_jump_hmc as S
    from tests.helpers import pin_init
    rng = np.random.default_rng(100 + d)
    np.random.seed(99)
    if dist_name == "Funnel":
        x0 = 3.0 * rng.standard_normal((1, n))
        X0 = np.concatenate((x0, np.exp(x0 / 2.) * rng.standard_normal((d - 1, n))))
        dist = pin_init(D.Funnel(scale=3.0, ndims=d, nbatch=n), X0)
        eps, L = 0.1, 7
    elif dist_name == "RoughWell":
        dist = pin_init(D.RoughWell(ndims=d, nbatch=n, scale1=5, scale2=2), 5 * rng.standard_normal((d, n)))
        eps, L = 0.3, 5
    else:
        dist = pin_init(D.Gaussian(ndims=d, nbatch=n, log_conditioning=2), rng.standard_normal((d, n)))
        eps, L = 0.4, 6
    kw = dict(resample=False) if kind in ("ContinuousTimeHMC", "MarkovJumpHMC") else {}
    s = getattr(S, kind)(distribution=dist, epsilon=eps, beta=0.3, num_leapfrog_steps=L, seed=3, dtype=dtype,
                         register_state=register_state, **kw)
    out = s.sample(n_iter)
    st = s.state
    counts = (s.l_count, s.f_count, s.fl_count, s.r_count, dist.E_count, dist.dEdX_count)
    extra = [np.array(s.dwelling_times)] if kind in ("ContinuousTimeHMC", "MarkovJumpHMC") else []
    return [np.asarray(out), np.array(st.X), np.array(st.V)] + extra, counts


@pytest.mark.parametrize("kind", ["HMCBase", "HMC", "ControlHMC", "ContinuousTimeHMC", "MarkovJumpHMC"])
@pytest.mark.parametrize("dist_name,d,dtype", [("Funnel", 10, "float64"), ("Funnel", 7, "float64"), ("Gaussian", 6, "float64"),
                                               ("RoughWell", 8, "float64"), ("Gaussian", 16, "float64"),
                                               ("RoughWell", 13, "float64"), ("Funnel", 10, "float32"),
                                               ("Gaussian", 16, "float32")])
def test_shared_memory_state_kernel_equals_the_register_kernel(dist_name, d, dtype, kind):
    a, ca = _stash_case(dist_name, d, dtype, kind, False)
    b, cb = _stash_case(dist_name, d, dtype, kind, True)
    assert ca == cb
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x.view(np.int64), y.view(np.int64))
