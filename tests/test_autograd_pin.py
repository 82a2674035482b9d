"""Independent pin for the two hand-derived gradients of the oracle.

The reference never writes these gradients down: ProductOfT asks Theano for ``T.grad(T.sum(energy), state)``
(misc/distributions.py:410) and the Funnel asks TensorFlow for ``tf.gradients(energy_op, state_pl)``
(misc/tf_distributions.py:89-91).  Neither framework is in this image, but reverse-mode differentiation of the
same expression graph is exactly what ``torch.autograd`` computes: the reference's energy formulas are written
here in torch operation by operation (misc/distributions.py:428-433, misc/tf_distributions.py:161-165) and
differentiated; the oracle's closed-form energies and gradients (oracle/mjhmc_oracle.py: ProductOfTEnergy,
FunnelEnergy) must agree to rounding.  CPU only, float64.
"""
import numpy as np
import pytest
import torch

from oracle import mjhmc_oracle as orc


def _pot_energy_torch(X, W, nu, b):
    """misc/distributions.py:428-433, line by line (X: [ndims, n])."""
    rshp_b = b.reshape((1, -1))
    rshp_nu = nu.reshape((1, -1))
    alpha = (rshp_nu + 1.) / 2.
    energy_per_expert = alpha * torch.log(1 + ((torch.matmul(X.T, W) + rshp_b) / rshp_nu) ** 2)
    return torch.sum(energy_per_expert, dim=1).reshape((1, -1))


def _funnel_literal_energy_torch(X, scale):
    """misc/tf_distributions.py:161-165 as written: e_x_0 [n] broadcasts over the ndims - 1 rows of e_x_k."""
    e_x_0 = -((X[0, :] ** 2) / (scale ** 2))
    e_x_k = -((X[1:, :] ** 2) / torch.exp(X[0, :]))
    return torch.sum(e_x_0 + e_x_k, dim=0)


def _funnel_neal_energy_torch(X, scale):
    """-log density of the distribution the reference docstring states (misc/tf_distributions.py:143-147):
    x_0 ~ N(0, scale^2), x_i ~ N(0, e^{x_0}) (variance), additive constants dropped."""
    d = X.shape[0]
    return X[0] ** 2 / (2 * scale ** 2) + 0.5 * torch.exp(-X[0]) * torch.sum(X[1:] ** 2, dim=0) + (d - 1) * X[0] / 2


def _autograd(fn, X):
    Xt = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    E = fn(Xt)
    G, = torch.autograd.grad(E.sum(), Xt)              # T.grad(T.sum(energy), state) / tf.gradients(energy, state)
    return E.detach().numpy().reshape(-1), G.numpy()


@pytest.mark.parametrize("d,kind", [(6, "identity"), (36, "sparse"), (100, "sparse"), (100, "dense")])
def test_product_of_t_gradient_is_the_autodiff_of_the_reference_energy(d, kind):
    rs = np.random.RandomState(2015 + d)
    if kind == "identity":
        W = np.eye(d)
    elif kind == "sparse":                             # search/MJHMC_poe_36/mjhmc_objective.py:15-23
        W = rs.randn(d, d)
        W[rs.rand(d, d) > 0.05] = 0
        W += np.eye(d) * 0.5
    else:
        W = rs.randn(d, d) / np.sqrt(d)
    W = W.astype(np.float32).astype(np.float64)        # parameters are float32 in the reference (:398-406)
    nu = (rs.rand(d) * 2 + 2.1).astype(np.float32).astype(np.float64)
    b = (rs.randn(d) * 0.3).astype(np.float32).astype(np.float64)
    X = rs.randn(d, 257) * 2.0
    en = orc.ProductOfTEnergy(W, nu, b)
    E_ref, G_ref = _autograd(lambda x: _pot_energy_torch(x, torch.tensor(W), torch.tensor(nu), torch.tensor(b)), X)
    np.testing.assert_allclose(np.asarray(en.E(X)).reshape(-1), E_ref, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(en.dEdX(X), G_ref, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("scale", [1.0, 3.0])
@pytest.mark.parametrize("d", [2, 10])
def test_funnel_gradients_are_the_autodiff_of_the_reference_graph(scale, d):
    rs = np.random.RandomState(7 * d)
    X = np.vstack((rs.randn(1, 300) * scale, rs.randn(d - 1, 300) * 2.0))
    lit = orc.FunnelEnergy(scale, literal=True)
    E_ref, G_ref = _autograd(lambda x: _funnel_literal_energy_torch(x, scale), X)
    np.testing.assert_allclose(np.asarray(lit.E(X)).reshape(-1), E_ref, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(lit.dEdX(X), G_ref, rtol=1e-12, atol=1e-12)
    neal = orc.FunnelEnergy(scale, literal=False)
    E_ref, G_ref = _autograd(lambda x: _funnel_neal_energy_torch(x, scale), X)
    np.testing.assert_allclose(np.asarray(neal.E(X)).reshape(-1), E_ref, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(neal.dEdX(X), G_ref, rtol=1e-12, atol=1e-12)


def _sparse_image_code_graph_torch(X, basis, patches, lmbda, cauchy, n_patches, n_coeffs):
    """misc/tf_distributions.py:240-268 op by op (tf.reshape is row-major like torch.reshape)."""
    n = X.shape[1]
    patches_r = patches.reshape(n_patches, 1, -1)
    shaped_state = X.reshape(n_patches, -1, n_coeffs, 1)
    shaped_basis = basis.reshape(1, 1, basis.shape[0], n_coeffs).expand(n_patches, n, -1, -1)
    reconstructions = torch.matmul(shaped_basis, shaped_state)[:, :, :, 0]
    reconstruction_error = torch.sum(0.5 * (patches_r - reconstructions) ** 2, -1)
    reconstruction_error = torch.mean(reconstruction_error, 0)
    if cauchy:
        sp_penalty = lmbda * torch.sum(torch.log(1 + X ** 2), 0)
    else:
        sp_penalty = lmbda * torch.sum(torch.abs(X), 0)
    return reconstruction_error + sp_penalty


@pytest.mark.parametrize("cauchy", [True, False])
@pytest.mark.parametrize("n", [1, 4])
def test_sparse_image_code_gradient_is_the_autodiff_of_the_reference_graph(cauchy, n):
    rs = np.random.RandomState(11)
    n_patches, n_coeffs, img = 3, 20, 12
    basis, patches = rs.randn(img, n_coeffs) / 3, rs.randn(n_patches, img)
    X = rs.randn(n_patches * n_coeffs, n)
    lit = orc.SparseImageCodeEnergy(basis, patches, 0.01, cauchy, literal=True)
    E_ref, G_ref = _autograd(lambda x: _sparse_image_code_graph_torch(x, torch.tensor(basis), torch.tensor(patches), 0.01,
                                                                      cauchy, n_patches, n_coeffs), X)
    np.testing.assert_allclose(np.asarray(lit.E(X)).reshape(-1), E_ref, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(lit.dEdX(X), G_ref, rtol=1e-12, atol=1e-13)
    # the particle-consistent form agrees with the graph for one particle and is column-separable for several
    fix = orc.SparseImageCodeEnergy(basis, patches, 0.01, cauchy, literal=False)
    if n == 1:
        np.testing.assert_allclose(fix.E(X), lit.E(X), rtol=1e-13)
        np.testing.assert_allclose(fix.dEdX(X), lit.dEdX(X), rtol=1e-12, atol=1e-13)
    else:
        for b in range(n):
            np.testing.assert_allclose(fix.E(X)[:, b], fix.E(X[:, b:b + 1])[:, 0], rtol=1e-13)
            np.testing.assert_allclose(fix.dEdX(X)[:, b], fix.dEdX(X[:, b:b + 1])[:, 0], rtol=1e-12, atol=1e-13)


def test_tf_fit_follows_autodiff_of_the_reference_graph_under_tf1_adam():
    """search/objective.py:137-183 builds ``curve = exp(a t) cos(b t)``, ``loss = reduce_sum((y - curve)^2)`` and lets
    ``tf.train.AdamOptimizer(learning_rate).minimize(loss)`` run.  TensorFlow is absent; the graph is written here in
    torch, differentiated by autograd, and stepped with the update rule the TF1 documentation of AdamOptimizer states
    (lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); m, v moving averages; var -= lr_t m / (sqrt(v) + epsilon); b1 = 0.9,
    b2 = 0.999, epsilon = 1e-8).  The package's tf_fit (numpy, analytic gradient) must walk the same path and return the
    parameters of the smallest loss seen."""
    from mjhmc_b200.search import objective
    rs = np.random.RandomState(0)
    t = np.linspace(0, 12, 60)
    y = np.exp(-0.35 * t) * np.cos(0.9 * t) + 0.01 * rs.randn(60)
    n_steps, lr = 300, 0.01
    a0, b0 = objective.estimate_params(t, y)
    params = torch.tensor([float(a0), float(b0)], dtype=torch.float64, requires_grad=True)
    tt, yt = torch.tensor(t), torch.tensor(y)
    m, v = torch.zeros(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
    losses, trail = [], []
    for step in range(1, n_steps + 1):
        curve = torch.exp(params[0] * tt) * torch.cos(params[1] * tt)
        loss = torch.sum((yt - curve) ** 2)
        g, = torch.autograd.grad(loss, params)
        losses.append(float(loss))
        trail.append(params.detach().clone().numpy())
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        lr_t = lr * np.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
        with torch.no_grad():
            params -= lr_t * m / (torch.sqrt(v) + 1e-8)
    want = trail[int(np.argmin(losses))]
    got = objective.tf_fit(t, y, n_steps=n_steps, learning_rate=lr)
    np.testing.assert_allclose(got, want, rtol=1e-10)
    assert min(losses) < losses[0]
