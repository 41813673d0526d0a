"""Fantasy models: the posterior conditioned on a *batch* of target draws at the same new inputs.

This is the shape BoTorch's ``fantasize`` produces (``online_gp/models/online_ski_botorch_model.py:51-61``:
``Y_fantasized`` is ``num_fantasies x q x t`` for ``X`` of shape ``q x d``).  The reference expands every cache —
including the dense m x m ``WtW`` and its root / inverse-root — once per fantasy (``get_fantasy_model``,
``batched_fixed_noise_online_gp.py:287-332``; ``UpdatedRootLazyTensor._expand_batch``,
``updated_root_lazy_tensor.py:139-159``).  Only the target-dependent caches differ between fantasies
(``interpolation_cache`` b_f = b + W^T D^-1 y_f, ``response_cache``); ``WtW``, hence L, B, K L, Q and the predictive
covariance, are shared.  So a fantasy batch is ONE conditioned model (panels updated once) plus an [m, nf] block of
interpolation caches pushed through the same kernels with nf right-hand sides:

    K b_f  : one Kronecker-Toeplitz MVM on an m x nf block        (reference: nf MVMs)
    c_f    : one Gram  L^T (K B)  -> r x nf                        (reference: nf passes over the m x r panel)
    mean_f : gathers of the stencil rows of K B and K L            (never the m-vector mu_u per fantasy)

``FantasizedOnlineSKIGP`` exposes what the acquisition code reads from such a model: ``eval()``, ``__call__(X)`` /
``posterior(X)`` with a leading fantasy dimension, ``num_fantasies``, ``num_data``.  Fantasy inputs that differ per
batch element (``X`` of shape ``b x q x d``) need one panel pair per element and are not covered here.
"""
import torch

from .. import ops
from ..distributions import MultivariateNormal
from ..lazy.lazy_tensor import LazyTensor, NonLazyTensor


class _SharedCovar(LazyTensor):
    """The same q x q covariance for every fantasy, presented with the fantasy batch shape."""

    def __init__(self, base, batch_shape):
        self.base, self._bs = base, torch.Size(batch_shape)

    def _size(self):
        return torch.Size((*self._bs, *self.base.shape))

    def evaluate(self):
        return self.base.evaluate().expand(*self._bs, *self.base.shape)

    def diag(self):
        d = self.base.diag()
        return d.expand(*self._bs, *d.shape)

    def _matmul(self, rhs):
        return self.evaluate() @ rhs

    def _transpose_nonbatch(self):
        return self

    dtype = property(lambda self: self.base.dtype)
    device = property(lambda self: self.base.device)


class FantasizedOnlineSKIGP:
    def __init__(self, conditioned, interpolation_cache, response_cache):
        """conditioned: the model conditioned on (X, any one target draw) — provides the shared operators;
        interpolation_cache [nf, t, m, 1], response_cache [nf, t, 1, 1]: the per-fantasy target caches."""
        self.model = conditioned
        self.interpolation_cache = interpolation_cache
        self.response_cache = response_cache
        self.num_fantasies = interpolation_cache.shape[0]
        self.num_outputs = conditioned.num_outputs
        self.num_data = conditioned.num_data
        self.covar_module, self.likelihood = conditioned.covar_module, conditioned.likelihood
        self._proj = None

    def eval(self):
        self.model.eval()
        return self

    def train(self, mode=True):
        raise RuntimeError("fantasy models are prediction-only")

    def _solve_caches(self):
        """Per output o: (K B_o [m, nf], a_o = Q^-1 L^T K B_o [r, nf]) for the block B_o of fantasy caches."""
        if self._proj is None:
            m = self.model
            out = []
            Ls = m._root_panels()
            with torch.no_grad():
                for o, (K, L, Q) in enumerate(zip(m.Kuu.items, Ls, m.current_qmatrix.items)):
                    Bo = self.interpolation_cache[:, o, :, 0].t().contiguous()               # [m, nf]
                    KB = K.detach()._matmul(Bo)
                    a = Q.inv_matmul(ops.gram(L, KB))            # Q's Cholesky factor is the conditioned model's cached one
                    out.append((KB, a))
            self._proj = out
        return self._proj

    def __call__(self, X):
        m = self.model
        if m.training:
            raise RuntimeError("call .eval() first: fantasy models are prediction-only")
        if X.dim() != 2:
            raise NotImplementedError("fantasy models take unbatched test inputs [q*, d]")
        base = m(X)                                            # shared covariance (+ builds the shared caches)
        lazy_kernel = m.covar_module(X).evaluate_kernel()
        idx, val = lazy_kernel.left_interp_indices, lazy_kernel.left_interp_values.detach()
        cache = m.prediction_cache
        means = []
        with torch.no_grad():
            for o, (KB, a) in enumerate(self._solve_caches()):
                T = ops.left_interp(idx, val, cache["KL"][o])                                  # [q*, r]
                means.append((ops.left_interp(idx, val, KB) - T @ a).t())                      # [nf, q*]
        mean = torch.stack(means, dim=1)                                                       # [nf, t, q*]
        cov = base.lazy_covariance_matrix
        if self.num_outputs == 1 and m._batch_shape == torch.Size():
            return MultivariateNormal(mean[:, 0], _SharedCovar(cov, (self.num_fantasies,)))
        return MultivariateNormal(mean, _SharedCovar(cov, (self.num_fantasies,)))

    forward = __call__

    def posterior(self, X, observation_noise=False, **kwargs):
        from .online_ski_botorch_model import GPyTorchPosterior
        self.eval()
        return GPyTorchPosterior(self(X))


def condition_on_target_batch(model, X, Y, noise):
    """``model.condition_on_observations(X, Y, noise)`` for Y of shape [nf, q, t] (noise [q, t] or [nf, q, t] with
    identical rows): returns a ``FantasizedOnlineSKIGP``."""
    nf, q, t = Y.shape
    if noise is None:
        noise = torch.ones(q, t, dtype=Y.dtype, device=Y.device)
    if noise.dim() == 3:
        if not bool((noise == noise[:1]).all()):
            raise NotImplementedError("fantasies with different noise per draw need one root update per draw")
        noise = noise[0]
    conditioned = model.condition_on_observations(X, Y[0], noise, inplace=False)
    lazy_kernel = model.covar_module(X).evaluate_kernel()
    idx, val = lazy_kernel.left_interp_indices, lazy_kernel.left_interp_values.detach()
    m = model.covar_module.num_inducing
    old = model._kernel_cache
    interp, resp = [], []
    for o in range(t):
        dinv_y = (Y[:, :, o] / noise[:, o]).t().contiguous()                                    # [q, nf]
        add = ops.left_t_interp(idx, val, dinv_y, m)                                            # [m, nf]
        interp.append((old["interpolation_cache"][o] + add).t().unsqueeze(-1))                  # [nf, m, 1]
        resp.append(old["response_cache"][o].reshape(1) + (Y[:, :, o] ** 2 / noise[:, o]).sum(-1))   # [nf]
    interpolation_cache = torch.stack(interp, dim=1)                                            # [nf, t, m, 1]
    response_cache = torch.stack(resp, dim=1).reshape(nf, t, 1, 1)
    fm = FantasizedOnlineSKIGP(conditioned, interpolation_cache, response_cache)
    if model.training is False:
        fm.eval()
    return fm
