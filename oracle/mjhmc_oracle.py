"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's sampler loop.

Header (task rule 3): this file is the *oracle*.  It is imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  The product (``mjhmc_b200``) never
imports it and has no CPU fallback.

What it restates (all file:line under /root/reference/mjhmc/):
  * leapfrog / L / F / FLF / R and the FLF cache  -- samplers/hmc_state.py:46-148
  * HMCBase / HMC / ControlHMC discrete chain      -- samplers/markov_jump_hmc.py:24-200
  * ContinuousTimeHMC jump process + resampling    -- samplers/markov_jump_hmc.py:203-347
  * MarkovJumpHMC incl. infinite-rate back-off     -- samplers/markov_jump_hmc.py:350-415
  * draw_from / min_idx                            -- misc/utils.py:15-49
  * energies: TestGaussian, Gaussian, RoughWell    -- misc/distributions.py:256-312,348-370
              ProductOfT (Theano graph restated)   -- misc/distributions.py:379-433
              Funnel (TensorFlow graph restated)   -- misc/tf_distributions.py:142-177
  * fft_autocor                                    -- misc/autocor.py:37-49

Pinning: for TestGaussian / Gaussian / RoughWell the oracle is pinned against
the reference itself run in the build container (tests/golden/*.npz, produced by
tests/golden/generate_golden.py from /root/reference) -- seeded ``np.random``
runs and injected-draw trajectories.  ProductOfT, Funnel and
fft_autocor cannot be run from the reference (Theano / TensorFlow / mklfft are
absent): for those **parity is unpinned**; their gradients are checked against
finite differences instead.

The state is kept as plain arrays (struct of arrays, ``(ndims, nbatch)`` float64
like the reference) and the FLF cache as one scalar + one flag per particle
(only ``H()`` of the cached state is ever read: markov_jump_hmc.py:367).
"""
import numpy as np

from . import philox as _philox


# ----------------------------------------------------------------------------
# energies
# ----------------------------------------------------------------------------
class Energy(object):
    """E(X) -> (n,), dEdX(X) -> (ndims, n) for X of shape (ndims, n)."""
    name = "energy"

    def E(self, X):
        raise NotImplementedError

    def dEdX(self, X):
        raise NotImplementedError


class TestGaussianEnergy(Energy):
    """distributions.py:357-362."""
    __test__ = False
    name = "TestGaussian"

    def __init__(self, sigma=1.):
        self.sigma = sigma

    def E(self, X):
        return np.sum(X ** 2, axis=0) / (2. * self.sigma ** 2)

    def dEdX(self, X):
        return X / self.sigma ** 2


class GaussianEnergy(Energy):
    """distributions.py:262-273 (J may be any square matrix; the reference builds a diagonal one)."""
    name = "Gaussian"

    def __init__(self, J):
        self.J = np.asarray(J, dtype=np.float64)

    @classmethod
    def log_conditioned(cls, ndims, log_conditioning=6):
        return cls(np.diag(10 ** np.linspace(-log_conditioning, 0, ndims)))

    def E(self, X):
        return np.sum(X * np.dot(self.J, X), axis=0) / 2.

    def dEdX(self, X):
        return np.dot(self.J, X) / 2. + np.dot(self.J.T, X) / 2.


class RoughWellEnergy(Energy):
    """distributions.py:295-304."""
    name = "RoughWell"

    def __init__(self, scale1=100, scale2=4):
        self.scale1 = scale1
        self.scale2 = scale2

    def E(self, X):
        cosX = np.cos(X * 2 * np.pi / self.scale2)
        return np.sum((X ** 2) / (2 * self.scale1 ** 2) + cosX, axis=0)

    def dEdX(self, X):
        sinX = np.sin(X * 2 * np.pi / self.scale2)
        return X / self.scale1 ** 2 + -sinX * 2 * np.pi / self.scale2


class MultimodalGaussianEnergy(Energy):
    """distributions.py:314-335: two unit Gaussians at -/+ sep_vec, sep_vec = (2 * separation, 0, ..., 0)
    (row 0 of the reference's array holds ``separation`` and gets ``+= separation``, :316-319)."""
    name = "MultimodalGaussian"

    def __init__(self, separation=3, ndims=2):
        self.separation = separation
        self.sep = np.zeros((ndims, 1))
        self.sep[0, 0] = 2 * separation

    def E(self, X):
        with np.errstate(over='ignore', divide='ignore'):
            return -np.log(np.exp(-np.sum((X + self.sep) ** 2, axis=0)) + np.exp(-np.sum((X - self.sep) ** 2, axis=0)))

    def dEdX(self, X):
        with np.errstate(over='ignore', invalid='ignore'):
            common_exp = np.exp(np.sum(4 * self.sep * X, axis=0))
            return (2 * ((X - self.sep) * common_exp + self.sep + X)) / (common_exp + 1)


class ProductOfTEnergy(Energy):
    """distributions.py:398-406 (float32 parameters) and :428-433 (energy).

    The gradient is the hand-derived autodiff of :431 (the reference asks Theano
    for it, :410):  G_j = (nu_j + 1) y_j / (nu_j**2 + y_j**2),  dEdX = W G^T.
    """
    name = "ProductOfT"

    def __init__(self, W, nu, b=None):
        self.W = np.array(W, dtype='float32').astype(np.float64)
        self.nu = np.array(nu, dtype='float32').astype(np.float64)
        nb = self.W.shape[1]
        self.b = (np.zeros(nb) if b is None else np.array(b, dtype='float32').astype(np.float64))

    def _Y(self, X):
        return np.dot(X.T, self.W) + self.b.reshape((1, -1))

    def E(self, X):
        nu = self.nu.reshape((1, -1))
        alpha = (nu + 1.) / 2.
        return np.sum(alpha * np.log(1 + (self._Y(X) / nu) ** 2), axis=1)

    def dEdX(self, X):
        nu = self.nu.reshape((1, -1))
        Y = self._Y(X)
        G = (nu + 1.) * Y / (nu ** 2 + Y ** 2)
        return np.dot(self.W, G.T)


class FunnelEnergy(Energy):
    """tf_distributions.py:158-165.

    literal=True follows the graph as written: e_x_0 (n,) broadcasts over the
    ndims-1 rows of e_x_k before the reduce_sum, i.e.
        E = -sum_{k>=1} [ x_0^2/scale^2 + x_k^2 exp(-x_0) ].
    literal=False (default of the build) is Neal's funnel that the
    docstring (:143-147) describes, x_0 ~ N(0, scale^2), x_k ~ N(0, e^{x_0}):
        E = x_0^2/(2 scale^2) + exp(-x_0)/2 sum_k x_k^2 + (ndims-1) x_0 / 2.
    """
    name = "Funnel"

    def __init__(self, scale=1.0, literal=False):
        self.scale = float(scale)
        self.literal = literal

    def E(self, X):
        x0, xk = X[0], X[1:]
        nk = xk.shape[0]
        if self.literal:
            e0 = -((x0 ** 2) / (self.scale ** 2))
            ek = -((xk ** 2) / np.exp(x0))
            return np.sum(e0 + ek, axis=0)
        return (x0 ** 2) / (2 * self.scale ** 2) + 0.5 * np.exp(-x0) * np.sum(xk ** 2, axis=0) + nk * x0 / 2.

    def dEdX(self, X):
        x0, xk = X[0], X[1:]
        nk = xk.shape[0]
        g = np.empty_like(X)
        s = np.sum(xk ** 2, axis=0)
        if self.literal:
            g[0] = -2. * nk * x0 / self.scale ** 2 + np.exp(-x0) * s
            g[1:] = -2. * xk * np.exp(-x0)
        else:
            g[0] = x0 / self.scale ** 2 - 0.5 * np.exp(-x0) * s + nk / 2.
            g[1:] = xk * np.exp(-x0)
        return g


class SparseImageCodeEnergy(Energy):
    """misc/tf_distributions.py:204-284 (SparseImageCode) restated in numpy.

    E = mean_p 1/2 ||patch_p - basis a_p||^2 + lmbda sum log(1 + x^2)   (:258-268; Laplace prior: sum |x|)
    literal=True keeps the reference's row-major reshape of the (ndims, nbatch) state to (n_patches, nbatch,
    n_coeffs) (:246); literal=False gives particle b the coefficient vectors a_p = x[p n_coeffs:(p+1) n_coeffs, b].
    The gradient is the hand-derived autodiff of the graph (pinned against torch.autograd in tests/test_autograd_pin.py).
    PARITY: TensorFlow is absent, the reference cannot run; pinned by autograd of the graph written op by op."""

    def __init__(self, basis, patches, lmbda=0.01, cauchy=True, literal=False):
        self.basis = np.asarray(basis, dtype=np.float64)          # [img_size, n_coeffs]
        self.patches = np.asarray(patches, dtype=np.float64)      # [n_patches, img_size]
        self.lmbda, self.cauchy, self.literal = lmbda, cauchy, literal
        self.n_patches, self.n_coeffs = self.patches.shape[0], self.basis.shape[1]

    def _shaped(self, X):
        n = X.shape[1]
        if self.literal:
            return X.reshape(self.n_patches, n, self.n_coeffs)
        return X.reshape(self.n_patches, self.n_coeffs, n).transpose(0, 2, 1)

    def E(self, X):
        resid = self.patches[:, None, :] - self._shaped(X) @ self.basis.T
        rec = np.mean(np.sum(0.5 * resid ** 2, axis=-1), axis=0)
        pen = np.sum(np.log(1 + X ** 2), axis=0) if self.cauchy else np.sum(np.abs(X), axis=0)
        return (rec + self.lmbda * pen).reshape(1, -1)

    def dEdX(self, X):
        n = X.shape[1]
        resid = self.patches[:, None, :] - self._shaped(X) @ self.basis.T
        G = -(resid @ self.basis) / self.n_patches
        G = G.reshape(X.shape[0], n) if self.literal else G.transpose(0, 2, 1).reshape(X.shape[0], n)
        return G + self.lmbda * (2 * X / (1 + X ** 2) if self.cauchy else np.sign(X))


class LambdaEnergy(Energy):
    """README.md:14-35 intent of LambdaDistribution (distributions.py:216-235)."""
    name = "Lambda"

    def __init__(self, energy_func, energy_grad_func):
        self._E = energy_func
        self._g = energy_grad_func

    def E(self, X):
        return np.asarray(self._E(X), dtype=np.float64).reshape(-1)

    def dEdX(self, X):
        return np.asarray(self._g(X), dtype=np.float64)


# ----------------------------------------------------------------------------
# draw sources
# ----------------------------------------------------------------------------
INFINITE_RATE_MSG = "Infinite rate"


def _exp_from_uniform(rates, u):
    """utils.py:38-42 with np.random.exponential(scale) == scale * -log(1 - u)."""
    rates = np.asarray(rates, dtype=np.float64)
    if not np.all(np.isfinite(rates)):
        raise ValueError(INFINITE_RATE_MSG)
    out = np.full(rates.shape, np.inf)
    nz = rates != 0
    out[nz] = (1. / rates[nz]) * (-np.log(1.0 - u[nz]))
    return out


class LegacyNumpyDraws(object):
    """Consumes the global ``np.random`` stream in exactly the reference's call order."""

    def normals(self, attempt, ndims, n):
        return np.random.randn(ndims, n)

    def uniforms(self, attempt, slot, n):
        return np.random.rand(n)

    def coin(self, attempt):
        return np.random.random()

    def exponentials(self, attempt, slot, rates):
        # utils.py:37-48: sequential; zero rates consume nothing; the first
        # non-finite rate raises after the draws before it were consumed.
        rates = np.asarray(rates, dtype=np.float64)
        bad = np.where(~np.isfinite(rates))[0]
        stop = bad[0] if len(bad) else len(rates)
        out = np.full(rates.shape, np.inf)
        nz = np.where(rates[:stop] != 0)[0]
        if len(nz):
            out[nz] = np.random.exponential(scale=1. / rates[nz])
        if len(bad):
            raise ValueError(INFINITE_RATE_MSG)
        return out

    def resample_uniforms(self, m):
        return np.random.random(m)


class InjectedDraws(object):
    """Pre-drawn arrays indexed by (attempt, slot, particle)  (SURVEY Appendix A.2).

    Z: (n_attempts, ndims, n)   normals for R
    U: (n_attempts, 3, n)       uniforms; slot meaning per sampler:
         MarkovJumpHMC      0:l  1:f  2:r      ContinuousTimeHMC 0:fl 1:f 2:r
         HMCBase/HMC/Control 0:accept 1:flip
    U0: (n_attempts,)           batch-wide R coin of the discrete samplers
    Ur: (m,)                    resampling uniforms
    """

    def __init__(self, Z, U, U0=None, Ur=None):
        self.Z, self.U, self.U0, self.Ur = Z, U, U0, Ur

    def normals(self, attempt, ndims, n):
        return np.array(self.Z[attempt], dtype=np.float64)

    def uniforms(self, attempt, slot, n):
        return np.asarray(self.U[attempt, slot], dtype=np.float64)

    def coin(self, attempt):
        return float(self.U0[attempt])

    def exponentials(self, attempt, slot, rates):
        return _exp_from_uniform(rates, self.uniforms(attempt, slot, len(rates)))

    def resample_uniforms(self, m):
        return np.asarray(self.Ur[:m], dtype=np.float64)


class PhiloxDraws(object):
    """The counter-based stream of the CUDA kernels (oracle/philox.py)."""

    def __init__(self, seed, particle_offset=0):
        self.seed = int(seed)
        self.offset = int(particle_offset)
        self._cache = (None, None)

    def _particles(self, n):
        return np.arange(self.offset, self.offset + n, dtype=np.uint64)

    def normals(self, attempt, ndims, n):
        return _philox.normals(self.seed, attempt, self._particles(n), ndims)

    def uniforms(self, attempt, slot, n):
        if self._cache[0] != (attempt, n):
            self._cache = ((attempt, n), _philox.uniforms(self.seed, attempt, self._particles(n)))
        return self._cache[1][slot]

    def coin(self, attempt):
        return _philox.coin(self.seed, attempt)

    def exponentials(self, attempt, slot, rates):
        return _exp_from_uniform(rates, self.uniforms(attempt, slot, len(rates)))

    def resample_uniforms(self, m):
        raise NotImplementedError("resampling uniforms are host-side draws")


# ----------------------------------------------------------------------------
# sampler
# ----------------------------------------------------------------------------
KINDS = ("HMCBase", "HMC", "ControlHMC", "ContinuousTimeHMC", "MarkovJumpHMC")


def derive_hyperparameters(kind, epsilon=1e-4, alpha=0.2, beta=None, num_leapfrog_steps=5):
    """markov_jump_hmc.py:67-80 (base), :189 (HMC), :197-200 (Control), :221-223 (CT/MJ).

    Returns dict(epsilon, beta, L, p_flip, p_r)."""
    beta = beta or alpha ** (1. / (epsilon * num_leapfrog_steps))
    p_flip, p_r = 0.5, 1
    if kind == "HMC":
        p_flip = 1
    elif kind in ("ControlHMC", "ContinuousTimeHMC", "MarkovJumpHMC"):
        if kind == "ControlHMC":
            p_flip = 1          # the jump samplers never read p_flip and keep the base value
        with np.errstate(divide='ignore'):
            p_r = - np.log(1 - beta) * 0.5
        beta = 1
    return dict(epsilon=epsilon, beta=beta, L=num_leapfrog_steps, p_flip=p_flip, p_r=p_r)


class OracleSampler(object):
    """One object for the five reference sampler classes (``kind``)."""

    def __init__(self, kind, energy, Xinit, V=None, epsilon=1e-4, alpha=0.2, beta=None,
                 num_leapfrog_steps=5, draws=None, resample=True, verbose=False):
        assert kind in KINDS
        self.kind = kind
        self.energy = energy
        self.draws = draws if draws is not None else LegacyNumpyDraws()
        self.verbose = verbose
        hp = derive_hyperparameters(kind, epsilon, alpha, beta, num_leapfrog_steps)
        self.epsilon, self.beta, self.num_leapfrog_steps = hp['epsilon'], hp['beta'], hp['L']
        self.p_flip, self.p_r = hp['p_flip'], hp['p_r']
        self.original_epsilon = epsilon
        self.resample = resample
        self.E_count = 0
        self.dEdX_count = 0
        self.l_count = self.f_count = self.fl_count = self.r_count = 0
        self.attempt = 0
        self.X = np.array(Xinit, dtype=np.float64)
        self.ndims, self.nbatch = self.X.shape
        if V is None:
            # hmc_state.py:26 -- drawn before the first energy evaluation
            V = self.draws.normals(-1, self.ndims, self.nbatch) if isinstance(self.draws, LegacyNumpyDraws) \
                else np.zeros_like(self.X)
        self.V = np.array(V, dtype=np.float64)
        # hmc_state.py:28-39
        self.EX = self._E(self.X)
        self.EV = self._kinetic(self.V)
        self.g = self._dEdX(self.X)
        # hmc_state.py:41-44, stored as H + flag
        self.cache_active = np.zeros(self.nbatch, dtype=bool)
        self.H_cache = np.zeros(self.nbatch)
        self.dwelling_times = np.zeros(self.nbatch)
        self.last_choice = np.full(self.nbatch, -1, dtype=np.int64)

    # -- counted evaluations (distributions.py:62-75) -------------------------
    def _E(self, X):
        self.E_count += X.shape[1]
        return np.asarray(self.energy.E(X), dtype=np.float64).reshape(-1)

    def _dEdX(self, X):
        self.dEdX_count += X.shape[1]
        return self.energy.dEdX(X)

    @staticmethod
    def _kinetic(V):
        return np.sum(V ** 2, axis=0) / 2.        # hmc_state.py:49-50

    def H(self):
        return self.EX + self.EV                  # hmc_state.py:80-84

    # -- hmc_state.py:86-100 --------------------------------------------------
    def _L(self, X, V, g):
        X, V, g = X.copy(), V.copy(), g.copy()
        for _ in range(self.num_leapfrog_steps):
            V += -self.epsilon / 2. * g
            X += self.epsilon * V
            g = self._dEdX(X)
            V += -self.epsilon / 2. * g
        EV = self._kinetic(V)
        EX = self._E(X)
        return X, V, g, EX, EV

    def _R(self, V, Z):
        # hmc_state.py:126-127
        return V * np.sqrt(1. - self.beta) + Z * np.sqrt(self.beta)

    # -- one iteration ----------------------------------------------------------
    def sampling_iteration(self):
        if self.kind == "MarkovJumpHMC":
            self._mj_iteration()
        elif self.kind == "ContinuousTimeHMC":
            self._ct_iteration()
        else:
            self._discrete_iteration()

    def _discrete_iteration(self):
        """markov_jump_hmc.py:116-148."""
        a = self.attempt
        self.attempt += 1
        N = self.nbatch
        Xp, Vp, gp, EXp, EVp = self._L(self.X, self.V, self.g)
        Vp = -Vp
        Ediff = self.H() - (EXp + EVp)
        p_acc = np.ones(N)
        with np.errstate(over='ignore', invalid='ignore'):
            neg = Ediff < 0
            p_acc[neg] = np.exp(Ediff[neg])
        acc = self.draws.uniforms(a, 0, N) < p_acc
        self._take(acc, Xp, Vp, gp, EXp, EVp)
        flip = self.draws.uniforms(a, 1, N) < self.p_flip * np.ones(N)
        self.V[:, flip] = -self.V[:, flip]
        if self.draws.coin(a) < self.p_r:
            self.r_count += N
            self.V = self._R(self.V, self.draws.normals(a, self.ndims, N))
            self.EV = self._kinetic(self.V)
        self.l_count += int(np.sum(acc & flip))
        self.f_count += int(np.sum(flip & ~acc))
        self.fl_count += int(np.sum(acc & ~flip))
        self.last_choice = np.where(acc, 0, 1)

    def _take(self, idx, X, V, g, EX, EV):
        self.X[:, idx] = X[:, idx]
        self.V[:, idx] = V[:, idx]
        self.g[:, idx] = g[:, idx]
        self.EX[idx] = EX[idx]
        self.EV[idx] = EV[idx]

    def _rates(self, H_to):
        # markov_jump_hmc.py:341-347
        with np.errstate(over='ignore', invalid='ignore'):
            return np.exp(self.H() - H_to) ** .5

    def _ct_iteration(self):
        """markov_jump_hmc.py:251-290."""
        a = self.attempt
        self.attempt += 1
        N = self.nbatch
        Xp, Vp, gp, EXp, EVp = self._L(self.X, self.V, self.g)
        Vp = -Vp
        fl_rates = self._rates(EXp + EVp)
        f_rates = np.ones(N)
        r_rates = self.p_r * np.ones(N)
        fl_draws = self.draws.exponentials(a, 0, fl_rates)
        f_draws = self.draws.exponentials(a, 1, f_rates)
        r_draws = self.draws.exponentials(a, 2, r_rates)
        stacked = np.stack([f_draws, fl_draws, r_draws])      # min_idx([f, fl, r]) :271
        choice = np.argmin(stacked, axis=0)
        self.dwelling_times = np.amin(stacked, axis=0)
        f_idx, fl_idx, r_idx = (choice == 0), (choice == 1), (choice == 2)
        self._take(fl_idx, Xp, Vp, gp, EXp, EVp)
        self.V[:, f_idx] = -self.V[:, f_idx]
        Vr = self._R(self.V, self.draws.normals(a, self.ndims, N))   # :285, after the exponentials
        self.V[:, r_idx] = Vr[:, r_idx]
        self.EV[r_idx] = self._kinetic(Vr)[r_idx]
        self.fl_count += int(fl_idx.sum())
        self.f_count += int(f_idx.sum())
        self.r_count += int(r_idx.sum())
        self.last_choice = choice

    def _mj_iteration(self):
        """markov_jump_hmc.py:355-415 incl. the back-off :376-389."""
        a = self.attempt
        self.attempt += 1
        N = self.nbatch
        # l_state :358
        Xl, Vl, gl, EXl, EVl = self._L(self.X, self.V, self.g)
        # flf_state :360 -> hmc_state.py:109-119 (only uncached particles integrate)
        act = ~self.cache_active
        H_flf = self.H_cache.copy()
        if act.any():
            _, _, _, EXf, EVf = self._L(self.X[:, act], -self.V[:, act], self.g[:, act])
            H_flf[act] = EXf + EVf
        # r_state :361 (normals drawn before the exponentials)
        Vr = self._R(self.V, self.draws.normals(a, self.ndims, N))
        EVr = self._kinetic(Vr)
        try:
            l_rates = self._rates(EXl + EVl)
            flf_rates = self._rates(H_flf)
            with np.errstate(invalid='ignore'):
                f_rates = flf_rates - np.min((flf_rates, l_rates), axis=0)
            r_rates = self.p_r * np.ones(N)
            l_draws = self.draws.exponentials(a, 0, l_rates)
            f_draws = self.draws.exponentials(a, 1, f_rates)
            r_draws = self.draws.exponentials(a, 2, r_rates)
        except ValueError:
            self.epsilon *= 0.5
            self.num_leapfrog_steps *= 2
            if self.verbose:
                depth = np.log(self.original_epsilon / self.epsilon) / np.log(2)
                print("Ecountered infinite rate, doubling back. Depth: {}".format(depth))
            self.cache_active[:] = False
            self._mj_iteration()
            self.epsilon *= 2
            self.num_leapfrog_steps = int(self.num_leapfrog_steps / 2)
            return
        stacked = np.stack([l_draws, f_draws, r_draws])       # min_idx([l, f, r]) :392
        choice = np.argmin(stacked, axis=0)
        self.dwelling_times = np.amin(stacked, axis=0)
        l_idx, f_idx, r_idx = (choice == 0), (choice == 1), (choice == 2)
        # :399 pre-transition state becomes the FLF state of the L movers
        self.H_cache[l_idx] = self.H()[l_idx]
        self.cache_active[l_idx] = True
        self._take(l_idx, Xl, Vl, gl, EXl, EVl)
        self.V[:, f_idx] = -self.V[:, f_idx]
        self.V[:, r_idx] = Vr[:, r_idx]
        self.EV[r_idx] = EVr[r_idx]
        self.cache_active[r_idx] = False
        self.cache_active[f_idx] = False
        self.l_count += int(l_idx.sum())
        self.f_count += int(f_idx.sum())
        self.r_count += int(r_idx.sum())
        self.last_choice = choice

    # -- sample() ---------------------------------------------------------------
    def burn_in(self, n_burn_in=500):
        for _ in range(n_burn_in):
            self.sampling_iteration()

    def sample(self, n_samples=1000, preserve_order=False):
        """markov_jump_hmc.py:150-173 (discrete) / :293-338 (CT, MJ)."""
        continuous = self.kind in ("ContinuousTimeHMC", "MarkovJumpHMC")
        if continuous and self.resample:
            samples_k, dwell_k = [], []
            self.sampling_iteration()
            samples_k.append(self.X.copy())
            for _ in range(n_samples):
                dwell_k.append(self.dwelling_times.copy())
                self.sampling_iteration()
                samples_k.append(self.X.copy())
            dwell_t = np.concatenate(dwell_k)
            samples = np.concatenate(samples_k[:-1], axis=1)
            r = np.sort(self.draws.resample_uniforms(n_samples * self.nbatch)) * np.sum(dwell_t)
            return np.ascontiguousarray(samples[:, resample_indices(dwell_t, r)])
        samples = []
        for _ in range(n_samples):
            self.sampling_iteration()
            samples.append(self.X.copy())
        if preserve_order:
            return np.stack(samples, axis=-1)
        return np.concatenate(samples, axis=1)

    def counters(self):
        return dict(l=self.l_count, f=self.f_count, fl=self.fl_count, r=self.r_count,
                    E=self.E_count, dEdX=self.dEdX_count)


def resample_indices(dwell_t, r):
    """markov_jump_hmc.py:324-328: first index with cumul_t > r (== searchsorted right)."""
    cumul_t = np.cumsum(dwell_t)
    return np.searchsorted(cumul_t, r, side='right')


# ----------------------------------------------------------------------------
# autocorrelation / ESS
# ----------------------------------------------------------------------------
def fft_autocor(samples):
    """autocor.py:37-49 with np.fft in place of mklfft: circular, no mean subtraction."""
    assert samples.ndim == 3
    f = np.fft.fft(samples, axis=-1)
    ac = np.real(np.mean(np.fft.ifft(f * np.conj(f), axis=-1), axis=(0, 1)))
    return ac / ac[0]


def ess_from_autocor(ac):
    """ESS = T / (1 + 2 sum_{tau>=1}^{first rho<0} rho_tau)  (defined by the build)."""
    T = len(ac)
    s = 0.0
    for tau in range(1, T):
        if ac[tau] < 0:
            break
        s += ac[tau]
    return T / (1.0 + 2.0 * s)


class FastNumpyDraws(object):
    """Vectorised draws from a numpy Generator: same distributions as the reference's
    per-particle ``np.random.exponential`` loop (utils.py:31-49) without the Python loop.
    Used only as the CPU-baseline configuration of bench.py (a faster, hence more
    conservative, baseline than the reference's own loop)."""

    def __init__(self, seed=0):
        self.g = np.random.default_rng(seed)

    def normals(self, attempt, ndims, n):
        return self.g.standard_normal((ndims, n))

    def uniforms(self, attempt, slot, n):
        return self.g.random(n)

    def coin(self, attempt):
        return float(self.g.random())

    def exponentials(self, attempt, slot, rates):
        return _exp_from_uniform(rates, self.g.random(len(rates)))

    def resample_uniforms(self, m):
        return self.g.random(m)
