"""CPU: the C-ABI library loads and exports every symbol include/mjhmc_b200.h declares, and the
host-side logic (hyper-parameter derivation, error behaviour, infinite-rate protocol, sample
layout) behaves like the reference -- exercised with a scripted stand-in for the device engine."""
import contextlib
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from mjhmc_b200 import _lib
from mjhmc_b200.misc import distributions as D
from mjhmc_b200.samplers import markov_jump_hmc as S
from oracle import mjhmc_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mjhmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mjhmc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libmjhmc_b200.so does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding and header disagree"


def test_host_only_entry_points():
    lib = _lib.load(build_if_missing=False)
    assert lib.mjhmc_abi_version() == _lib.ABI_VERSION == 2
    assert lib.mjhmc_resample_scratch_bytes(0) >= 0
    assert lib.mjhmc_resample_scratch_bytes(10 ** 6) >= 8 * 10 ** 6
    d = _lib.Dist()
    d.kind, d.dtype, d.ndims = _lib.DIST_ROUGH_WELL, _lib.F64, 2
    assert lib.mjhmc_fused_supported(ctypes.byref(d)) == 1
    d.ndims = 16
    assert lib.mjhmc_fused_supported(ctypes.byref(d)) == 1
    d.ndims = 17                                    # above the register kernel: the streaming kernel (<= 128 dims)
    assert lib.mjhmc_fused_supported(ctypes.byref(d)) == 1 and lib.mjhmc_stream_supported(ctypes.byref(d)) == 1
    d.ndims = 129
    assert lib.mjhmc_fused_supported(ctypes.byref(d)) == 0 and lib.mjhmc_stream_supported(ctypes.byref(d)) == 0
    d.kind, d.ndims = _lib.DIST_FUNNEL, 20          # not separable: no streaming kernel
    assert lib.mjhmc_fused_supported(ctypes.byref(d)) == 0 and lib.mjhmc_stream_supported(ctypes.byref(d)) == 0
    d.kind = 99
    assert lib.mjhmc_fused_supported(ctypes.byref(d)) == 0
    assert b"bad distribution kind" in lib.mjhmc_last_error()
    # argument checking happens before any device work
    d.kind, d.ndims = _lib.DIST_ROUGH_WELL, 2
    assert lib.mjhmc_energy(ctypes.byref(d), None, 5, 5, None, None) == -1
    assert b"NULL" in lib.mjhmc_last_error()


def test_product_path_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.ControlHMC(distribution=D.RoughWell(2, 10))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D.RoughWell(2, 10).E(np.zeros((2, 3)))


def test_markov_jump_requires_distribution():
    """markov_jump_hmc.py:236-242."""
    X = np.zeros((2, 4))
    with pytest.raises(NotImplementedError):
        S.MarkovJumpHMC(X, lambda x: x, lambda x: x)


# ---------------------------------------------------------------------------------------------
class FakeEngine(object):
    """Scripted device engine: records launches, returns counters from a script."""
    script = []          # list of fail iterations (None = no failure), consumed per launch
    log = []

    def __init__(self, sampler, distribution, X0, V0, opts):
        self.sampler = sampler
        self.fused = True
        self.device = torch.device("cpu")
        self.tdtype = torch.float64
        self.d, self.n = X0.shape
        self.dwell_last = torch.zeros(self.n, dtype=torch.float64)
        self.launches = 0
        self.commits = 0
        self.cache_resets = 0

    def ctx(self):
        return contextlib.nullcontext()

    def launch(self, attempt0, n_iter, samples=None, it0=0, dwell=None, choice=None, energy=None):
        s = self.sampler
        fail = FakeEngine.script.pop(0) if FakeEngine.script else None
        FakeEngine.log.append(dict(attempt0=attempt0, n_iter=n_iter, it0=it0, eps=s.epsilon, L=s.num_leapfrog_steps,
                                   fail=fail))
        self.launches += 1
        cnt = [0] * _lib.N_COUNTERS
        cnt[_lib.CNT_FAIL] = _lib.INT64_MAX if fail is None else fail
        cnt[_lib.CNT_L] = n_iter * self.n
        cnt[_lib.CNT_E] = n_iter * self.n
        cnt[_lib.CNT_DEDX] = n_iter * self.n * s.num_leapfrog_steps
        if samples is not None and fail is None:
            for j in range(n_iter):
                samples[:, it0 + j, :] = float(attempt0 + j) + torch.arange(self.n, dtype=torch.float64) / 1000.0
        return cnt

    def commit(self):
        self.commits += 1

    def reset_cache(self):
        self.cache_resets += 1

    def download(self):
        return np.zeros((self.d, self.n)), np.zeros((self.d, self.n)), np.zeros(self.n, bool), np.zeros(self.n)

    def upload(self, st):
        pass


@pytest.fixture
def fake_engine(monkeypatch):
    monkeypatch.setattr(S, "_Engine", FakeEngine)
    FakeEngine.script, FakeEngine.log = [], []
    return FakeEngine


KINDS = dict(HMCBase=S.HMCBase, HMC=S.HMC, ControlHMC=S.ControlHMC, ContinuousTimeHMC=S.ContinuousTimeHMC,
             MarkovJumpHMC=S.MarkovJumpHMC)


@pytest.mark.parametrize("kind", sorted(KINDS))
@pytest.mark.parametrize("hp", [dict(), dict(epsilon=0.3, beta=0.25, num_leapfrog_steps=7),
                                dict(epsilon=0.5, alpha=0.4, num_leapfrog_steps=3)])
def test_hyperparameter_derivation_matches_reference(fake_engine, kind, hp):
    """markov_jump_hmc.py:67-80,189,197-200,221-223 (oracle.derive_hyperparameters restates them)."""
    s = KINDS[kind](distribution=D.TestGaussian(2, 6), **hp)
    want = orc.derive_hyperparameters(kind, **hp)
    assert (s.epsilon, s.beta, s.num_leapfrog_steps, s.p_flip) == (want["epsilon"], want["beta"], want["L"], want["p_flip"])
    assert s.p_r == want["p_r"]
    assert s.n_burn_in == 500 and s.grad_per_sample_step == s.num_leapfrog_steps
    assert (s.ndims, s.nbatch) == (2, 6)
    assert s.original_epsilon == s.epsilon and s.original_l == s.num_leapfrog_steps


def test_misspelt_kwarg_is_a_type_error(fake_engine):
    """tests/test_continuous_samplers.py:69-85 of the reference passes num_leapfropg_steps -> TypeError."""
    with pytest.raises(TypeError):
        S.ControlHMC(distribution=D.Gaussian(), beta=0.3, epsilon=1.0, num_leapfropg_steps=3)


def test_constructor_counters_and_init_draw_order(fake_engine):
    """A.2 / A.3: reset() regenerates Xinit, the state constructor counts N energy and N gradient
    evaluations; the non-MJ ContinuousTimeHMC builds its state twice and ends at N as well (Q18)."""
    for kind, n_randn in (("ControlHMC", 2), ("MarkovJumpHMC", 2), ("ContinuousTimeHMC", 4)):
        np.random.seed(5)
        dist = D.TestGaussian(3, 11)                      # randn #0 (constructor)
        KINDS[kind](distribution=dist)
        assert (dist.E_count, dist.dEdX_count) == (11, 11), kind
        after = np.random.rand()
        np.random.seed(5)
        for _ in range(n_randn + 1):
            np.random.randn(3, 11)
        assert np.random.rand() == after, kind


def test_sample_layout(fake_engine):
    s = S.ControlHMC(distribution=D.TestGaussian(2, 5))
    X = s.sample(3)
    assert X.shape == (2, 15) and X.dtype == np.float64
    # time-major, particle-minor columns (np.concatenate(samples, axis=1), markov_jump_hmc.py:173)
    np.testing.assert_allclose(X[0], np.concatenate([t + np.arange(5) / 1000.0 for t in range(3)]))
    Y = s.sample(num_steps=2, preserve_order=True)            # README.md:35 alias
    assert Y.shape == (2, 5, 2)
    np.testing.assert_allclose(Y[1, :, 1], 4 + np.arange(5) / 1000.0)
    assert s.l_count == 5 * 5 and s.distribution.dEdX_count == 5 + 5 * 5 * s.num_leapfrog_steps


def test_markov_jump_backoff_protocol(fake_engine):
    """markov_jump_hmc.py:376-389: failure at iteration 2 of 5 -> replay 2, count the failed attempt's
    evaluations, retry that iteration at eps/2, 2L with a wiped cache, restore, continue."""
    dist = D.TestGaussian(2, 4)
    s = S.MarkovJumpHMC(distribution=dist, epsilon=1.0, beta=0.5, num_leapfrog_steps=3, resample=False)
    e0, g0 = dist.E_count, dist.dEdX_count
    FakeEngine.script[:] = [2, None, 0, None, None]
    s.sample(5)
    log = FakeEngine.log
    assert [(l["attempt0"], l["n_iter"], l["it0"], l["eps"], l["L"]) for l in log] == [
        (0, 5, 0, 1.0, 3),        # fails at relative iteration 2
        (0, 2, 0, 1.0, 3),        # replay of the two good iterations
        (2, 1, 0, 1.0, 3),        # the failed attempt (evaluations counted, state discarded)
        (3, 1, 2, 0.5, 6),        # retry with halved step, doubled steps
        (4, 2, 3, 1.0, 3),        # the rest with the restored hyper-parameters
    ]
    assert (s.epsilon, s.num_leapfrog_steps) == (1.0, 3)
    assert s._engine.cache_resets == 1 and s._engine.commits == 3
    assert s._attempt == 6
    # l counts only for committed launches; E/dEdX also for the failed attempt
    assert s.l_count == (2 + 1 + 2) * 4
    assert dist.E_count - e0 == (2 + 1 + 1 + 2) * 4
    assert dist.dEdX_count - g0 == (2 * 3 + 1 * 3 + 1 * 6 + 2 * 3) * 4


def test_continuous_time_raises_value_error(fake_engine):
    """No handler in ContinuousTimeHMC: the ValueError of draw_from reaches the caller (utils.py:41-48)."""
    dist = D.TestGaussian(2, 4)
    s = S.ContinuousTimeHMC(distribution=dist, epsilon=1.0, beta=0.5, num_leapfrog_steps=3, resample=False)
    FakeEngine.script[:] = [1, None, 0]
    e0 = dist.E_count
    with pytest.raises(ValueError, match="Infinite rate"):
        s.sample(4)
    assert s._engine.commits == 1 and dist.E_count - e0 == (1 + 1) * 4


def test_distribution_contract(fake_engine):
    g = D.Gaussian(ndims=4, nbatch=7, log_conditioning=2)
    np.testing.assert_allclose(np.diag(g.J), 10 ** np.linspace(-2, 0, 4))
    assert g.Xinit.shape == (4, 7) and g.E_count == 0 and g.backend == 'numpy'
    assert hash(g) == hash(D.Gaussian(ndims=4, nbatch=9, log_conditioning=2))       # nbatch not in the hash
    r = D.RoughWell(ndims=3, nbatch=5)
    assert (r.scale1, r.scale2) == (100, 4) and hash(r) == hash((3, 100, 4))
    f = D.Funnel(scale=2.0, nbatch=8, ndims=5)
    assert f.Xinit.shape == (5, 8) and hash(f) == hash((2.0, 5))
    with pytest.raises(NotImplementedError):
        D.ProductOfT(ndims=4, nbasis=5)
    p = D.ProductOfT(ndims=4, nbasis=4, nbatch=6)
    assert p.weights.dtype == np.float32 and p.Xinit.shape == (4, 6)
    lam = D.LambdaDistribution(energy_func=lambda X: np.sum(X ** 2, axis=0) / 2., energy_grad_func=lambda X: X,
                               init=np.ones((2, 3)), name="iso")
    assert lam.E(np.ones((2, 3))).shape == (1, 3) and lam.E_count == 3
    np.testing.assert_array_equal(lam.dEdX(np.ones((2, 3))), np.ones((2, 3)))
    assert lam.reset() is lam and lam.E_count == 0
    rot = D.Gaussian.rotated(ndims=5, nbatch=4, log_conditioning=2, seed=1)
    assert not rot._diagonal and np.allclose(rot.J, rot.J.T)
    np.testing.assert_allclose(np.sort(np.linalg.eigvalsh(rot.J)), 10 ** np.linspace(-2, 0, 5))


def test_install_alias(fake_engine):
    import mjhmc_b200
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == "mjhmc" or k.startswith("mjhmc.")}
    try:
        mjhmc_b200.install_alias()
        from mjhmc.samplers.markov_jump_hmc import MarkovJumpHMC
        from mjhmc.misc.distributions import LambdaDistribution
        from mjhmc.misc.tf_distributions import Funnel
        assert MarkovJumpHMC is S.MarkovJumpHMC and LambdaDistribution is D.LambdaDistribution and Funnel is D.Funnel
        assert MarkovJumpHMC.__name__ == "MarkovJumpHMC"       # string-compared by callers (autocor.py:29)
    finally:
        for k in [k for k in sys.modules if k == "mjhmc" or k.startswith("mjhmc.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_every_bench_workload_has_a_cpu_arm():
    """bench.py --impl reference / cpu_baseline: the oracle energy, the synthetic cloud and one sampling iteration of
    every workload (tiny particle count), so a new workload cannot ship without its CPU arm."""
    import bench
    from oracle import mjhmc_oracle as orc
    for name, w in bench.WORKLOADS.items():
        X, V = bench._init_cloud(w, 6, 0)
        assert X.shape == V.shape == (w["ndims"], 6), name
        s = orc.OracleSampler(w["sampler"], bench._oracle_energy(w), X, V=V, epsilon=w["epsilon"], beta=w["beta"],
                              num_leapfrog_steps=min(w["L"], 2), draws=orc.FastNumpyDraws(1), resample=False)
        s.sampling_iteration()
        assert np.all(np.isfinite(s.X)), name
        assert bench.algorithmic_bytes_per_launch(w) > 0


def test_ess_definition():
    import bench
    ac = np.array([1.0, 0.5, 0.25, -0.1, 0.3])
    assert abs(bench._ess_from_curve(ac, 100) - 100 / (1 + 2 * 0.75)) < 1e-12
    assert bench._ess_from_curve(np.array([1.0, -0.2]), 10) == 10.0
