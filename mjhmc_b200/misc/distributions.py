"""Distribution contract and the built-in energies (reference: mjhmc/misc/distributions.py,
Funnel from mjhmc/misc/tf_distributions.py:142-177).

Same public surface as the reference -- ``E(X)`` / ``dEdX(X)`` over ``(ndims, n)`` arrays with
``E_count`` / ``dEdX_count`` incremented by the number of columns, ``Xinit``, ``reset()``,
``gen_init_X()`` ... -- but the built-ins evaluate on the GPU through libmjhmc_b200 and carry a
*kernel descriptor* so the samplers can run them inside the fused sampler kernel.  Anything
without a descriptor (LambdaDistribution, user subclasses) runs in the unfused callback mode.

Differences from the reference, all listed in SURVEY.md Appendix A.4:
  Q4  the 1e6-step fair-initialisation burn-in behind every constructor is opt-in; ``init_X``
      uses ``gen_init_X``
  Q3  LambdaDistribution calls the lambdas it is given
  Q16 Funnel defaults to Neal's funnel; ``literal_reference_energy=True`` gives the TF graph as written
"""
import numpy as np
import torch

from .. import _device, _lib
from .utils import overrides


class Distribution(object):
    """Interface/abstract class for distributions (reference distributions.py:13-195)."""

    def __init__(self, ndims=2, nbatch=100):
        self.ndims = ndims
        self.nbatch = nbatch
        if not hasattr(self, 'backend'):
            self.backend = 'numpy'
        # true iff being sampled with a jump process
        self.mjhmc = None
        self.E_count = 0
        self.dEdX_count = 0
        self.generation_instance = False
        if not hasattr(self, 'max_n_particles'):
            self.max_n_particles = None
        self.init_X()

    # -- counted evaluation (distributions.py:62-75) -------------------------------------
    def E(self, X):
        self.E_count += X.shape[1]
        return self.E_val(X)

    def E_val(self, X):
        raise NotImplementedError()

    def dEdX(self, X):
        self.dEdX_count += X.shape[1]
        return self.dEdX_val(X)

    def dEdX_val(self, X):
        raise NotImplementedError()

    def __hash__(self):
        raise NotImplementedError()

    # -- initialisation --------------------------------------------------------------------
    def init_X(self):
        """Sets self.Xinit.  The reference routes this through a burn-in cache
        (distributions.py:96-149); here that is opt-in."""
        try:
            self.gen_init_X()
        except NotImplementedError:
            self.Xinit = np.random.randn(self.ndims, self.nbatch)   # distributions.py:139

    def gen_init_X(self):
        raise NotImplementedError()

    def reset(self):
        """resets the object. returns self for convenience (distributions.py:162-170)"""
        self.E_count = 0
        self.dEdX_count = 0
        if not self.generation_instance:
            self.init_X()
        return self

    def __call__(self, X):
        """NUTS convenience: returns -E, -dEdX (distributions.py:172-180)."""
        rshp_X = X.reshape(len(X), 1)
        E = float(np.asarray(self.E(rshp_X)).reshape(-1)[0])
        dEdX = np.asarray(self.dEdX(rshp_X)).T[0]
        return -E, -dEdX

    # -- B200 extension: which fused kernel evaluates this energy ---------------------------
    def kernel_descriptor(self, dtype, device):
        """None -> unfused callback mode.  Built-ins return a filled ``_lib.Dist`` plus the
        device tensors it points into (kept alive by the caller)."""
        return None


class _DeviceEnergy(Distribution):
    """Built-in energies: E_val / dEdX_val run the stand-alone device kernels."""
    # the unfused sampler path may hand device tensors straight to E / dEdX
    accepts_device_arrays = True

    def _eval(self, X, want_grad):
        lib = _lib.load()
        if isinstance(X, torch.Tensor):
            dev = _device.require_cuda(X.device)
            dtype = _device.norm_dtype(X.dtype)
            Xd = X.contiguous()
        else:
            dev = _device.require_cuda()
            dtype = "float64"
            Xd = _device.to_device(np.asarray(X, dtype=np.float64), dtype, dev)
        d, n = Xd.shape
        assert d == self.ndims, "X must be (ndims, n)"
        desc, keep = self.kernel_descriptor(dtype, dev)
        out = torch.empty((d, n) if want_grad else (1, n), dtype=Xd.dtype, device=dev)
        fn = lib.mjhmc_gradient if want_grad else lib.mjhmc_energy
        _lib.check(fn(desc, _device.ptr(Xd), n, n, _device.ptr(out), _device.stream_ptr(dev)),
                   "gradient" if want_grad else "energy")
        del keep
        return out if isinstance(X, torch.Tensor) else out.cpu().numpy()

    @overrides(Distribution)
    def E_val(self, X):
        return self._eval(X, False)

    @overrides(Distribution)
    def dEdX_val(self, X):
        return self._eval(X, True)

    def _desc(self, kind, dtype, p=(), arrays=(), nbasis=0):
        desc = _lib.Dist()
        desc.kind = kind
        desc.dtype = _device.dtype_code(dtype)
        desc.ndims = self.ndims
        desc.nbasis = nbasis
        for i, v in enumerate(p):
            desc.p[i] = float(v)
        ptrs = [a.data_ptr() if a is not None else 0 for a in arrays] + [0, 0, 0]
        desc.a0, desc.a1, desc.a2 = ptrs[0] or None, ptrs[1] or None, ptrs[2] or None
        return desc, tuple(arrays)


class LambdaDistribution(Distribution):
    """Anonymous distribution from an energy and a gradient callable (README.md:14-35,
    distributions.py:198-251).  Runs in the unfused callback mode: the state lives on the GPU,
    the callables are invoked on host arrays between the leapfrog kernels."""

    def __init__(self, energy_func=None, energy_grad_func=None, init=None, name=None):
        self.energy_func = energy_func
        self.energy_grad_func = energy_grad_func
        self.init = init
        self.name = name or str(np.random.random())
        super(LambdaDistribution, self).__init__(ndims=init.shape[0], nbatch=init.shape[1])

    @overrides(Distribution)
    def E_val(self, X):
        return np.asarray(self.energy_func(X)).reshape((1, -1))

    @overrides(Distribution)
    def dEdX_val(self, X):
        return np.asarray(self.energy_grad_func(X))

    @overrides(Distribution)
    def gen_init_X(self):
        self.Xinit = self.init

    @overrides(Distribution)
    def __hash__(self):
        return hash((self.ndims, self.name))


class Gaussian(_DeviceEnergy):
    """Ill-conditioned Gaussian of the LAHMC paper (distributions.py:256-281).

    ``J`` (B200 extension) replaces the reference's diagonal ``J`` by any square matrix, e.g. a
    rotated full-covariance precision; the reference code path ``np.dot(J, X)`` is already dense
    (distributions.py:268-273), only its constructor never builds a non-diagonal ``J``."""

    def __init__(self, ndims=2, nbatch=100, log_conditioning=6, J=None):
        self.conditioning = 10 ** np.linspace(-log_conditioning, 0, ndims)
        if J is None:
            self.J = np.diag(self.conditioning)
            self._diagonal = True
        else:
            self.J = np.array(J, dtype=np.float64)
            assert self.J.shape == (ndims, ndims)
            self._diagonal = bool(np.all(self.J == np.diag(np.diag(self.J))))
        self.description = '%dD Anisotropic Gaussian, %g self.conditioning' % (ndims, 10 ** log_conditioning)
        super(Gaussian, self).__init__(ndims, nbatch)

    @classmethod
    def rotated(cls, ndims=2, nbatch=100, log_conditioning=6, seed=0):
        """Full-covariance variant J = Q^T diag(cond) Q, Q from the QR of RandomState(seed).randn."""
        cond = 10 ** np.linspace(-log_conditioning, 0, ndims)
        Q, _ = np.linalg.qr(np.random.RandomState(seed).randn(ndims, ndims))
        return cls(ndims, nbatch, log_conditioning, J=Q.T.dot(np.diag(cond)).dot(Q))

    @overrides(Distribution)
    def gen_init_X(self):
        if self._diagonal:
            self.Xinit = (1. / np.sqrt(np.diag(self.J)).reshape((-1, 1))) * np.random.randn(self.ndims, self.nbatch)
        else:
            w, Q = np.linalg.eigh((self.J + self.J.T) / 2.)
            self.Xinit = Q.dot((1. / np.sqrt(w)).reshape((-1, 1)) * np.random.randn(self.ndims, self.nbatch))

    @overrides(Distribution)
    def __hash__(self):
        if self._diagonal:
            return hash((self.ndims, hash(tuple(np.diag(self.J)))))
        return hash((self.ndims, hash(tuple(self.J.ravel()))))

    def kernel_descriptor(self, dtype, device):
        if self._diagonal and self.ndims <= 16:
            j = _device.to_device(np.diag(self.J), dtype, device)
            return self._desc(_lib.DIST_DIAG_GAUSSIAN, dtype, arrays=(j,))
        S = _device.to_device((self.J + self.J.T) / 2., dtype, device)      # dEdX = J X/2 + J^T X/2
        if _device.norm_dtype(dtype) == "float32" and device is not None:
            # tcgen05 path: the matrix pre-tiled and split into tf32 hi / lo parts, prepared once
            lib = _lib.load()
            ws = torch.empty(int(lib.mjhmc_dense_tf32_workspace_bytes(self.ndims)), dtype=torch.uint8, device=device)
            desc, keep = self._desc(_lib.DIST_DENSE_GAUSSIAN, dtype, arrays=(S, ws))
            _lib.check(lib.mjhmc_dense_tf32_prepare(desc, _device.stream_ptr(device)), "dense_tf32_prepare")
            return desc, keep
        return self._desc(_lib.DIST_DENSE_GAUSSIAN, dtype, arrays=(S,))


class RoughWell(_DeviceEnergy):
    """Rough well of the LAHMC paper (distributions.py:283-312)."""

    def __init__(self, ndims=2, nbatch=100, scale1=100, scale2=4):
        self.scale1 = scale1
        self.scale2 = scale2
        self.description = '{} Rough Well'.format(ndims)
        super(RoughWell, self).__init__(ndims, nbatch)

    @overrides(Distribution)
    def gen_init_X(self):
        self.Xinit = self.scale1 * np.random.randn(self.ndims, self.nbatch)

    @overrides(Distribution)
    def __hash__(self):
        return hash((self.ndims, self.scale1, self.scale2))

    def kernel_descriptor(self, dtype, device):
        return self._desc(_lib.DIST_ROUGH_WELL, dtype, p=(self.scale1, self.scale2))


class TestGaussian(_DeviceEnergy):
    """Unit-variance Gaussian for testing samplers (distributions.py:348-370)."""
    __test__ = False

    def __init__(self, ndims=2, nbatch=100, sigma=1.):
        self.sigma = sigma
        super(TestGaussian, self).__init__(ndims, nbatch)

    @overrides(Distribution)
    def gen_init_X(self):
        self.Xinit = np.random.randn(self.ndims, self.nbatch)

    @overrides(Distribution)
    def __hash__(self):
        return hash((self.ndims, self.sigma))

    def kernel_descriptor(self, dtype, device):
        return self._desc(_lib.DIST_TEST_GAUSSIAN, dtype, p=(self.sigma,))


class ProductOfT(_DeviceEnergy):
    """Product of Student-t experts (distributions.py:373-453).  Parameters are rounded to
    float32 like the reference's Theano shared variables (:398-406); the gradient is the
    hand-derived autodiff of :431."""

    def __init__(self, ndims=36, nbasis=36, nbatch=100, lognu=None, W=None, b=None):
        if ndims != nbasis:
            raise NotImplementedError("Initializer only works for ndims == nbasis")
        self.ndims = ndims
        self.nbasis = nbasis
        self.nbatch = nbatch
        if W is None:
            W = np.eye(ndims, nbasis)
        self.weights = np.array(W, dtype='float32')
        if lognu is None:
            pre_nu = np.random.rand(nbasis,) * 2 + 2.1
        else:
            pre_nu = np.exp(lognu)
        self.nu = np.array(pre_nu, dtype='float32')
        if b is None:
            b = np.zeros((nbasis,))
        self.bias = np.array(b, dtype='float32')
        super(ProductOfT, self).__init__(ndims, nbatch)
        self.backend = 'cuda'

    @overrides(Distribution)
    def gen_init_X(self):
        from scipy import stats
        Zinit = np.zeros((self.ndims, self.nbatch))
        for ii in range(self.ndims):
            Zinit[ii] = stats.t.rvs(self.nu[ii], size=self.nbatch)
        Yinit = Zinit - self.bias.reshape((-1, 1))
        self.Xinit = np.dot(np.linalg.inv(self.weights), Yinit)

    @overrides(Distribution)
    def __hash__(self):
        return hash((self.ndims, self.nbasis, hash(tuple(self.nu)), hash(tuple(self.weights.ravel())),
                     hash(tuple(self.bias.ravel()))))

    def kernel_descriptor(self, dtype, device):
        W = _device.to_device(self.weights.astype(np.float64), dtype, device)
        nu = _device.to_device(self.nu.astype(np.float64), dtype, device)
        b = _device.to_device(self.bias.astype(np.float64), dtype, device)
        return self._desc(_lib.DIST_PRODUCT_OF_T, dtype, arrays=(W, nu, b), nbasis=self.nbasis)


class Funnel(_DeviceEnergy):
    """Neal's funnel (tf_distributions.py:142-177): x_0 ~ N(0, scale^2), x_i ~ N(0, e^{x_0}).

    The default energy is the distribution the reference docstring describes; the TensorFlow
    graph the reference actually builds (:158-165) is sign-flipped and unnormalised and is
    available as ``literal_reference_energy=True`` for short-horizon parity."""

    def __init__(self, scale=1.0, nbatch=50, ndims=10, literal_reference_energy=False, **kwargs):
        self.scale = float(scale)
        self.literal_reference_energy = literal_reference_energy
        self.name = 'Funnel'
        super(Funnel, self).__init__(ndims, nbatch)
        self.backend = 'cuda'

    @overrides(Distribution)
    def gen_init_X(self):
        x_0 = np.random.normal(scale=self.scale, size=(1, self.nbatch))
        x_k = np.random.normal(scale=np.exp(x_0), size=(self.ndims - 1, self.nbatch))
        self.Xinit = np.vstack((x_0, x_k))

    @overrides(Distribution)
    def __hash__(self):
        return hash((self.scale, self.ndims))

    def kernel_descriptor(self, dtype, device):
        kind = _lib.DIST_FUNNEL_LITERAL if self.literal_reference_energy else _lib.DIST_FUNNEL
        return self._desc(kind, dtype, p=(self.scale,))
