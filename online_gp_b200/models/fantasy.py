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


# ------------------------------------------------------------------------------------------------ batched fantasy inputs
def posterior_cross_cov(model, o, idx_a, val_a, idx_b, val_b):
    """Latent posterior covariance  W_a M W_b^T * scale  between two stencil sets for output o ([qa, qb]), with
    M = K - K L Q^-1 L^T K the grid-space predictive covariance (``_make_predictive_covar``, :385-404) and
    scale = sigma^2 when the model has the learnable noise factor (:227-228).  Differentiable w.r.t. the
    interpolation values (acquisition optimisation back-propagates into the candidate inputs)."""
    from ..lazy.lazy_tensor import _scatter_dense
    m = model.covar_module.num_inducing
    cache = model.prediction_cache
    K, KL, Q = model.Kuu.items[o].detach(), cache["KL"][o].detach(), model.current_qmatrix.items[o]
    eye = torch.eye(idx_b.shape[0], dtype=val_b.dtype, device=val_b.device)
    grad = torch.is_grad_enabled() and val_b.requires_grad
    Wb_t = _scatter_dense(idx_b, val_b, eye, m) if grad else ops.left_t_interp(idx_b, val_b, eye, m)      # [m, qb]
    c1 = ops.left_interp(idx_a, val_a, K._matmul(Wb_t))                                                    # [qa, qb]
    Ta = ops.left_interp(idx_a, val_a, KL)                                                                 # [qa, r]
    Tb = ops.left_interp(idx_b, val_b, KL)                                                                 # [qb, r]
    Lq = Q.cholesky().detach()
    cov = c1 - Ta @ torch.cholesky_solve(Tb.t(), Lq)
    if model.has_learnable_noise:
        cov = cov * model._second_noise(o).detach()
    return cov


class PredictiveSpaceFantasy:
    """Fantasy models for candidate batches X [b, q, d] that differ per batch element (BoTorch ``fantasize`` inside
    ``optimize_acqf``: look-ahead acquisition functions such as qNIPV, ``experiments/active_learning``).

    The reference gives every batch element its own copy of the dense ``WtW`` and of both root panels and updates each
    (``_expand_batch`` repeats them, ``updated_root_lazy_tensor.py:139-159``).  A fantasy is only ever *queried* —
    posterior mean / variance at test points — so it is represented here in predictive space instead: with the
    current posterior's latent covariance C(.,.) (``posterior_cross_cov``) and S_b = C(X_b, X_b) + sigma^2 D_b,

        mean_b(x*) = mean(x*) + C(x*, X_b) S_b^-1 (y_b - mean(X_b)),   var_b(x*) = C(x*, x*) - C(x*, X_b) S_b^-1 C(X_b, x*).

    This is the exact Gaussian conditional under the SKI kernel; it coincides with conditioning the WISKI caches
    whenever the root is exact (r = m: every configuration below ``max_cholesky_size``, i.e. all BO / active-learning
    setups of the reference), and no m-sized state is copied.  Y may be None (variance-only acquisition functions)."""

    def __init__(self, model, X, Y=None, noise=None):
        if X.dim() != 3:
            raise ValueError("X must be [b, q, d]")
        self.model = model.eval()
        self.b, self.q, _ = X.shape
        self.t = model._num_models
        self._X_flat = X.reshape(-1, X.shape[-1])
        lazy_kernel = model.covar_module(self._X_flat).evaluate_kernel()
        self.idx_b, self.val_b = lazy_kernel.left_interp_indices, lazy_kernel.left_interp_values
        if noise is None:
            noise = torch.ones(self.b, self.q, self.t, dtype=X.dtype, device=X.device)
        self.noise = noise.expand(self.b, self.q, self.t)
        self.Y = None if Y is None else (Y if Y.dim() == 4 else Y.unsqueeze(0))          # [nf, b, q, t]
        self.num_fantasies = None if self.Y is None else self.Y.shape[0]
        self.num_data = model.num_data + self.q
        self.num_outputs = model.num_outputs
        self._chol = {}

    def eval(self):
        return self

    def _solve_setup(self, o):
        """Cholesky factors of S_b = C(X_b, X_b) + sigma^2 D_b for every batch element: [b, q, q]."""
        if o not in self._chol:
            m = self.model
            C = posterior_cross_cov(m, o, self.idx_b, self.val_b, self.idx_b, self.val_b)            # [bq, bq]
            blocks = torch.stack([C[i * self.q:(i + 1) * self.q, i * self.q:(i + 1) * self.q] for i in range(self.b)])
            obs = self.noise[..., o]
            if m.has_learnable_noise:
                obs = obs * m._second_noise(o).detach()
            self._chol[o] = torch.linalg.cholesky(blocks + torch.diag_embed(obs))
        return self._chol[o]

    def __call__(self, X_test):
        """X_test [q*, d] -> MultivariateNormal-like with mean [(nf,) b, (t,) q*] and variance of the same trailing
        shape (diagonal only: look-ahead acquisition functions read ``.variance`` / ``.mean``)."""
        m = self.model
        if X_test.dim() != 2:
            raise NotImplementedError("test inputs must be [q*, d]")
        lk = m.covar_module(X_test).evaluate_kernel()
        idx_s, val_s = lk.left_interp_indices, lk.left_interp_values.detach()
        base = m(X_test)
        base_mean = base.mean.reshape(self.t, -1)
        base_var = base.variance.reshape(self.t, -1)
        base_at_b = m(self.idx_to_inputs()) if self.Y is not None else None
        means, variances = [], []
        for o in range(self.t):
            Lb = self._solve_setup(o)                                                               # [b, q, q]
            Csb = posterior_cross_cov(m, o, idx_s, val_s, self.idx_b, self.val_b)                   # [q*, bq]
            Csb = Csb.reshape(-1, self.b, self.q).permute(1, 0, 2)                                  # [b, q*, q]
            gain_t = torch.cholesky_solve(Csb.transpose(-1, -2), Lb)                                # [b, q, q*] = S^-1 C(X_b, x*)
            variances.append(base_var[o].unsqueeze(0) - (Csb.transpose(-1, -2) * gain_t).sum(-2))   # [b, q*]
            if self.Y is not None:
                resid = self.Y[..., o] - base_at_b.mean.reshape(self.t, self.b, self.q)[o]          # [nf, b, q]
                means.append(base_mean[o] + torch.einsum("fbq,bqs->fbs", resid, gain_t))            # [nf, b, q*]
        var = torch.stack(variances, dim=1)                                                         # [b, t, q*]
        if self.Y is None:
            mean = base_mean.unsqueeze(0).expand(self.b, self.t, -1)
        else:
            mean = torch.stack(means, dim=2)                                                        # [nf, b, t, q*]
            var = var.unsqueeze(0).expand(self.num_fantasies, *var.shape)
        if self.t == 1:
            mean, var = mean.squeeze(-2), var.squeeze(-2)
        return _DiagNormal(mean, var.clamp_min(0.0))

    def idx_to_inputs(self):
        return self._X_flat

    def posterior(self, X, observation_noise=False, **kwargs):
        from .online_ski_botorch_model import GPyTorchPosterior
        return GPyTorchPosterior(self(X))


class _DiagNormal:
    """Mean / variance container with the members ``GPyTorchPosterior`` reads."""

    def __init__(self, mean, variance):
        self.mean, self.variance = mean, variance

    def rsample(self, sample_shape=torch.Size(), base_samples=None):
        eps = torch.randn(*sample_shape, *self.mean.shape, dtype=self.mean.dtype, device=self.mean.device) \
            if base_samples is None else base_samples
        return self.mean + self.variance.sqrt() * eps
