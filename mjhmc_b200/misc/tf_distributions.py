"""The distributions the reference builds as TensorFlow graphs (reference: mjhmc/misc/tf_distributions.py).

``Funnel`` (:142-177) has a fused register kernel and lives in ``distributions``; it is re-exported here under the
reference's module path.  ``SparseImageCode`` (:204-284) is a sparse-coding posterior over
``n_patches * n_coeffs`` coefficients (9216 at the reference's defaults) for at most 50 particles: far outside the
shape of the fused kernels (<= 128 dims, millions of particles), so it runs through the UNFUSED device path of the
samplers -- state, leapfrog pieces and transition are this package's kernels, the energy and its gradient are the
batched ``[img x coeff]`` contraction below evaluated on device tensors (cuBLAS through torch.matmul: a plain library
GEMM, which is what it is) between them.  TensorFlow session / graph / profiling plumbing (:21-140) and ``TFGaussian``
(:179-202, a second statement of the unit Gaussian) are out of scope.
"""
import os
import pickle

import numpy as np
import torch

from .. import _device
from .distributions import Distribution, Funnel  # noqa: F401  (Funnel: reference module path)
from .utils import overrides


class SparseImageCode(Distribution):
    """ Distribution over the coefficients in an inference model of sparse coding on natural images a la Olshausen
    and Field (tf_distributions.py:204-284).

    E(x) = mean_p 1/2 || patch_p - basis . a_p ||^2 + lmbda * sum log(1 + x^2)      (Cauchy prior; Laplace: sum |x|)

    ``literal_reference_graph``: the reference reshapes the (ndims, nbatch) placeholder straight to
    (n_patches, nbatch, n_coeffs) (:246), which pairs coefficients of DIFFERENT particles in one reconstruction
    whenever nbatch > 1 (a row-major reshape, not a transpose).  False (default) gives every particle its own
    coefficient vectors a_p = x[p * n_coeffs : (p + 1) * n_coeffs]; True reproduces the graph as written.
    The two agree for nbatch == 1, the only setting the reference's experiments use (experiments/spectral.py:245).

    The image / basis blobs of the reference (``distr_data/dump_<n_basis>.pkl``) are absent from its repository
    (.MISSING_LARGE_BLOBS); pass ``data=dict(data=imgs [img_size, n_imgs], basis=[img_size, n_coeffs])`` or
    ``synthetic=True`` for seeded random blobs of the same shapes.
    """
    accepts_device_arrays = True

    def __init__(self, n_patches=9, n_batches=10, cauchy=True, n_basis=1024, data=None, synthetic=False,
                 literal_reference_graph=False, img_size=256, **kwargs):
        self.max_n_particles = 50
        self.lmbda = 0.01
        if data is None:
            assert n_basis in [1024, 512] or synthetic
            from .utils import package_path
            data_path = os.path.join(package_path(), "distr_data", "dump_{}.pkl".format(n_basis))
            if os.path.exists(data_path):
                with open(data_path, 'rb') as dump_file:
                    data = pickle.load(dump_file)
            elif synthetic:
                rs = np.random.RandomState(n_basis)
                data = dict(data=rs.randn(img_size, max(64, n_patches)), basis=rs.randn(img_size, n_basis) / np.sqrt(img_size))
            else:
                raise IOError("{} is missing (the reference ships no image blobs): pass data=... or synthetic=True"
                              .format(data_path))
        # [img_size, n_imgs]
        self.imgs = np.asarray(data['data'], dtype=np.float64)
        # [img_size, n_coeffs]
        self.basis = np.asarray(data['basis'], dtype=np.float64)
        self.img_size, self.n_coeffs = self.basis.shape
        self.n_patches = n_patches
        self.cauchy = cauchy
        self.literal_reference_graph = literal_reference_graph
        # [n_patches, img_size]
        self.patches = self.imgs[:, :n_patches].T.copy()
        self.name = 'SparseImageCode'
        self._dev = {}
        super(SparseImageCode, self).__init__(ndims=n_patches * self.n_coeffs, nbatch=n_batches)
        self.backend = 'cuda'

    # -- device evaluation ---------------------------------------------------------------------------------------
    def _consts(self, device, dtype):
        key = (str(device), dtype)
        if key not in self._dev:
            self._dev = {key: (torch.as_tensor(self.basis, device=device, dtype=dtype),
                               torch.as_tensor(self.patches, device=device, dtype=dtype))}
        return self._dev[key]

    def _shaped(self, Xd):
        n = Xd.shape[1]
        if self.literal_reference_graph:                       # tf.reshape(state_pl, [n_patches, -1, n_coeffs, 1]) :246
            return Xd.reshape(self.n_patches, n, self.n_coeffs)
        return Xd.reshape(self.n_patches, self.n_coeffs, n).permute(0, 2, 1)

    def _unshaped(self, G, n):
        if self.literal_reference_graph:
            return G.reshape(self.ndims, n)
        return G.permute(0, 2, 1).reshape(self.ndims, n)

    def _eval(self, X, want_grad):
        host = not isinstance(X, torch.Tensor)
        dev = _device.require_cuda(None if host else X.device)
        Xd = torch.as_tensor(np.asarray(X, dtype=np.float64), device=dev) if host else X
        basis, patches = self._consts(dev, Xd.dtype)
        n = Xd.shape[1]
        A = self._shaped(Xd)                                   # [n_patches, n, n_coeffs]
        resid = patches[:, None, :] - torch.matmul(A, basis.T)  # [n_patches, n, img_size]  (tf.batch_matmul :258)
        if not want_grad:
            rec = (0.5 * resid * resid).sum(dim=-1).mean(dim=0)                     # :260-262
            pen = torch.log1p(Xd * Xd).sum(dim=0) if self.cauchy else Xd.abs().sum(dim=0)
            out = (rec + self.lmbda * pen).reshape(1, n)
        else:
            G = self._unshaped(-torch.matmul(resid, basis) / self.n_patches, n)     # d/da of the reconstruction term
            out = G + self.lmbda * (2 * Xd / (1 + Xd * Xd) if self.cauchy else torch.sign(Xd))
        return out.cpu().numpy() if host else out

    @overrides(Distribution)
    def E_val(self, X):
        return self._eval(X, False)

    @overrides(Distribution)
    def dEdX_val(self, X):
        return self._eval(X, True)

    @overrides(Distribution)
    def __hash__(self):
        # (tf_distributions.py:276-284: bytes of the blobs and the scalars)
        return hash((hash(self.imgs.tobytes()), hash(self.basis.tobytes()), hash(self.lmbda), hash(self.n_patches),
                     self.n_coeffs))
