"""``FixedNoiseOnlineSKIGP`` — the WISKI model core (constant-in-n caches, Woodbury posterior, conditioning).

Same public surface as ``online_gp/models/batched_fixed_noise_online_gp.py`` (constructor ``:64-76`` incl. the
``kernel_cache=`` + ``num_data=`` re-hydration form, ``forward`` ``:173-256``, ``condition_on_observations``
``:258-285``, ``get_fantasy_model`` ``:287-332``, cached properties ``:334-383``, ``_make_predictive_covar``
``:385-404``, ``_dump_caches`` / ``zero_grad`` / ``set_train_data`` / ``to`` ``:406-435``) with the arithmetic
re-designed matrix-free on CUDA:

  * W^T is never densified (``_get_wmat_from_kernel`` ``:22-28`` is kept for API parity only): cache accumulation
    is a stencil scatter, the projection ``B^T v`` a row gather of the inverse-root panel;
  * ``_kernel_cache["WtW"]`` is an ``UpdatedRootLazyTensor`` holding the m x r root / inverse-root panels; the
    dense m x m matrix exists only in the Cholesky regime (m <= max_cholesky_size);
  * ``pred_cov`` (``K - K L Q^-1 L^T K``) stays an operator instead of the reference's dense m x m product.

Initial root beyond the Cholesky regime (the reference leaves this to GPyTorch's randomly started Lanczos,
SURVEY.md §7-H2): with V1 = W1^T D1^-1/2 of the first n1 = min(n0, max_root_decomposition_size) points and
G = V1^T V1 = U diag(lam) U^T, keep lam_j > tol lam_max:  L = V1 U,  B = L diag(1/lam); remaining initial points
are folded in with the projected update — all at once (``UpdatedRootLazyTensor.fold_in_sparse``: the point-by-point
projected updates telescope to one r x r factor and one panel GEMM per panel).  The CPU oracle
``oracle/wiski_matfree.py`` states the same rule point by point.
"""
import torch
from torch import nn

from .. import ops, settings
from ..distributions import MultivariateNormal
from ..kernels import GridInterpolationKernel, RBFKernel, ScaleKernel, _PriorMixin
from ..lazy.lazy_tensor import (BatchLazyTensor, InterpolatedLazyTensor, KroneckerToeplitzLazyTensor, LazyTensor,
                                NonLazyTensor, PanelLazyTensor, RootLazyTensor, ZeroLazyTensor, _scatter_dense)
from ..lazy.updated_root_lazy_tensor import UpdatedRootLazyTensor
from ..likelihoods import FNMGLikelihood
from ..settings import (detach_interp_coeff, fast_pred_samples, fast_pred_var, max_cholesky_size, overlap_root_update,
                        skip_posterior_variances)
from ..utils.memoize import CachingError, cached, pop_from_cache


def _get_wmat_from_kernel(lazy_kernel):
    """Dense W^T (m x q) — reference helper (``:22-28``); NOT used by the hot path of this implementation."""
    wmat = lazy_kernel._sparse_left_interp_t(lazy_kernel.left_interp_indices, lazy_kernel.left_interp_values).to_dense()
    if detach_interp_coeff.on():
        wmat = wmat.detach()
    return wmat


def _stencils(lazy_kernel):
    idx, val = lazy_kernel.left_interp_indices, lazy_kernel.left_interp_values
    if detach_interp_coeff.on():
        val = val.detach()
    return idx, val


def _wt_scatter(idx, val, src, m):
    """W^T src (m x c).  When the interpolation values carry a graph (features of a trainable stem inside ``fit()``, the
    reference differentiates its dense W^T unless ``detach_interp_coeff`` is on, ``:22-28``) the scatter is the
    autograd-capable one; otherwise the CUDA scatter kernel."""
    if torch.is_grad_enabled() and val.requires_grad:
        return _scatter_dense(idx, val, src, m)
    return ops.left_t_interp(idx, val, src, m)


def _lowrank_initial_roots(idx, vval, m, max_rank):
    """Deterministic initial (root, inv_root) beyond the Cholesky regime — see the module docstring."""
    n0 = idx.shape[0]
    n1 = min(n0, max_rank)
    dtype, device = vval.dtype, vval.device
    V1 = _wt_scatter(idx[:n1], vval[:n1], torch.eye(n1, dtype=dtype, device=device), m)
    lam, U = torch.linalg.eigh(ops.gram(V1, V1))
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    keep = lam.detach() > tol * lam.detach().max()
    lam, U = lam[keep].flip(0), U[:, keep].flip(1)
    r_eff = lam.numel()
    r = ((r_eff + 15) // 16) * 16                      # zero columns stay zero under every update
    Upad = torch.nn.functional.pad(U, (0, r - r_eff))
    scale = torch.nn.functional.pad(1.0 / lam, (0, r - r_eff))
    L = ops.panel_rmul(V1, Upad)
    B = (L * scale).contiguous()
    return L, B, n1


def _initialize_caches(targets, noise_diagonal, stencils, m, create_w_cache=True):
    """``_initialize_caches`` (``:31-60``) from stencils.  targets [n,t]; noise_diagonal [t,n]; returns the reference's
    cache dict: response_cache [t,1,1], interpolation_cache [t,m,1], D_logdet [t], WtW (batched operator)."""
    idx, val = stencils
    if targets.dim() == 1:
        targets = targets.unsqueeze(-1)
    y = targets.transpose(-1, -2)                       # [t,n]
    if noise_diagonal.dim() > 2:
        noise_diagonal = noise_diagonal.squeeze(-1)
    noise_diagonal = noise_diagonal.expand_as(y)
    dinv_y = y / noise_diagonal                         # :42
    cache = {
        "response_cache": (y * dinv_y).sum(-1).reshape(-1, 1, 1),                                        # :45
        "interpolation_cache": _wt_scatter(idx, val, dinv_y.t().contiguous(), m).t().unsqueeze(-1).contiguous(),  # :46
    }
    if create_w_cache:                                  # :49-53
        t, n = y.shape
        vvals = val.unsqueeze(0) / noise_diagonal.clamp_min(1e-7).sqrt().unsqueeze(-1)      # [t,n,s]
        if m <= settings.max_cholesky_size.value():
            tens = []
            for o in range(t):
                Vt = _wt_scatter(idx, vvals[o].contiguous(), torch.eye(n, dtype=val.dtype, device=val.device), m)
                tens.append(Vt @ Vt.t())
            cache["WtW"] = UpdatedRootLazyTensor(torch.stack(tens), initial_is_root=False)
        else:
            roots, invs = [], []
            n1 = n
            for o in range(t):
                L, B, n1 = _lowrank_initial_roots(idx, vvals[o].contiguous(), m, settings.max_root_decomposition_size.value())
                roots.append(L)
                invs.append(B)
            rmax = max(L.shape[1] for L in roots)
            pad = lambda P: torch.nn.functional.pad(P, (0, rmax - P.shape[1]))
            wtw = UpdatedRootLazyTensor(None, initial_is_root=False, root=torch.stack([pad(L) for L in roots]),
                                        inv_root=torch.stack([pad(B) for B in invs]))
            if n > n1:          # initial points beyond the root rank: one batched projected update (fold_in_sparse)
                wtw.fold_in_sparse(idx[n1:], vvals[:, n1:].contiguous())
            cache["WtW"] = wtw
    cache["D_logdet"] = noise_diagonal.log().sum(-1)    # :55
    return cache


class WoodburyInnerCovar(LazyTensor):
    """M = K - K L Q^-1 L^T K  (``_make_predictive_covar``, ``:385-404``) kept as an operator over the grid."""

    def __init__(self, Kuu, KL, qmatrix):
        self.Kuu, self.KL, self.qmatrix = Kuu, KL, qmatrix

    def _size(self):
        return self.Kuu.shape

    def _matmul(self, rhs):
        return self.Kuu._matmul(rhs) - ops.panel_rmul(self.KL, self.qmatrix.inv_matmul(ops.gram(self.KL, rhs)))

    def _transpose_nonbatch(self):
        return self

    def detach(self):
        return WoodburyInnerCovar(self.Kuu.detach(), self.KL.detach(), self.qmatrix.detach())

    def root_decomposition(self, method=None):
        raise NotImplementedError("root of the grid-space predictive covariance is not materialised (m x m)")

    dtype = property(lambda self: self.KL.dtype)
    device = property(lambda self: self.KL.device)


class PredictiveCovar(LazyTensor):
    """W* M W*^T * scale for q* test points (eval ``forward``, ``:222-228``): only q* x q* is ever formed."""

    def __init__(self, idx, val, inner, scale=None, T=None):
        self.idx, self.val, self.inner, self.scale, self.T = idx, val, inner, scale, T

    def _size(self):
        q = self.idx.shape[0]
        return torch.Size((q, q))

    def _block(self, sl):
        idx, val = self.idx[sl], self.val[sl]
        m = self.inner.shape[-1]
        eye = torch.eye(idx.shape[0], dtype=val.dtype, device=val.device)
        Wt = _scatter_dense(idx, val, eye, m) if val.requires_grad else ops.left_t_interp(idx, val, eye, m)
        c1 = ops.left_interp(idx, val, self.inner.Kuu._matmul(Wt))
        if self.T is not None and sl == slice(None):
            T = self.T
        else:
            T = ops.left_interp(idx, val, self.inner.KL).transpose(-1, -2)   # (K L)^T W*^T : r x q by row gather
        c2 = T.transpose(-1, -2) @ self.inner.qmatrix.inv_matmul(T)
        cov = c1 - c2
        return cov if self.scale is None else cov * self.scale

    def evaluate(self):
        if not hasattr(self, "_eval"):
            self._eval = self._block(slice(None))
        return self._eval

    def diag(self):
        q = self.idx.shape[0]
        if q <= 256 or hasattr(self, "_eval"):
            return self.evaluate().diagonal()
        return torch.cat([self._block(slice(s, s + 256)).diagonal() for s in range(0, q, 256)])

    def _matmul(self, rhs):
        return self.evaluate() @ rhs

    def _transpose_nonbatch(self):
        return self

    dtype = property(lambda self: self.val.dtype)
    device = property(lambda self: self.val.device)


class _PredictionCache(dict):
    """``prediction_cache`` dict (``:368-383``) whose "pred_mean" entry — the m-vector K b - K L Q^-1 c — is only
    materialised when somebody asks for it (one pass over the K L panel)."""

    def __init__(self, Kuu_response, KLs, qmat_solve):
        super().__init__()
        dict.__setitem__(self, "KL", KLs)
        dict.__setitem__(self, "qmat_solve", qmat_solve)
        self._Kuu_response = Kuu_response

    def has_pred_mean(self):
        return dict.__contains__(self, "pred_mean")

    def __getitem__(self, key):
        if key == "pred_mean" and not dict.__contains__(self, key):
            dict.__setitem__(self, key, self._Kuu_response - torch.stack(
                [ops.panel_rmul(KL, s) for KL, s in zip(dict.__getitem__(self, "KL"), dict.__getitem__(self, "qmat_solve"))]))
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        return key == "pred_mean" or dict.__contains__(self, key)

    def keys(self):
        return list(dict.keys(self)) + ([] if self.has_pred_mean() else ["pred_mean"])


class GP(nn.Module, _PriorMixin):
    pass


class FixedNoiseOnlineSKIGP(GP):
    def __init__(
        self,
        train_inputs=None,
        train_targets=None,
        train_noise_term=None,
        covar_module=None,
        kernel_cache=None,
        grid_bounds=None,
        grid_size=30,
        likelihood=None,
        learn_additional_noise=False,
        num_data=None,
        per_output_noise=False,
    ):
        super().__init__()

        assert train_inputs is not None or kernel_cache is not None

        if train_targets is not None:
            num_outputs = train_targets.shape[-1]
            input_batch_shape = train_inputs.shape[:-2]
            self.num_data = train_inputs.shape[-2]
        else:
            # pull from kernel_cache (``:86-90``)
            num_outputs = kernel_cache["response_cache"].shape[-1]
            input_batch_shape = kernel_cache["WtW"].shape[0]
            self.num_data = num_data

        self.num_outputs = num_outputs

        _batch_shape = input_batch_shape
        if num_outputs > 1:
            _batch_shape += torch.Size([num_outputs])

        if covar_module is None:
            if grid_bounds is None:
                grid_bounds = torch.stack(
                    (train_inputs.min(dim=-2)[0] - 0.1, train_inputs.max(dim=-2)[0] + 0.1)
                ).transpose(-1, -2)
            covar_module = ScaleKernel(
                RBFKernel(batch_shape=_batch_shape, ard_num_dims=train_inputs.size(-1)),
                batch_shape=_batch_shape,
            )

        if type(covar_module) is not GridInterpolationKernel:
            covar_module = GridInterpolationKernel(
                base_kernel=covar_module,
                grid_size=grid_size,
                num_dims=train_inputs.shape[-1],
                grid_bounds=grid_bounds,
            )

        self._batch_shape = _batch_shape
        self._num_data_t = None          # optional device-side copy of ``num_data`` (OnlineSKIRegression graph mode)
        self.train_inputs = [None]
        self.train_targets = None

        self.covar_module = covar_module
        if likelihood is None:
            if train_noise_term is None:
                train_noise_term = torch.ones_like(train_targets)
            # the learnable multiplicative noise is ONE scalar shared by all outputs, as in the reference (``:127-130``
            # builds the likelihood without a batch shape, so ``raw_noise`` has shape [1] and a reference state_dict
            # loads); ``per_output_noise=True`` (an addition) gives every output its own sigma^2 instead
            noise_t = train_noise_term.transpose(-1, -2)
            self.likelihood = FNMGLikelihood(
                noise=noise_t,
                learn_additional_noise=learn_additional_noise,
                batch_shape=noise_t.shape[:-1] if per_output_noise else torch.Size(),
            )
        else:
            self.likelihood = likelihood
        self.has_learnable_noise = learn_additional_noise

        # the training data are folded into the four caches here and not kept (``:92-95``)
        if kernel_cache is None:
            self.covar_module = self.covar_module.to(train_inputs.device)
            self.likelihood = self.likelihood.to(train_inputs.device)
            initial_kxx = self.covar_module(train_inputs).evaluate_kernel()
            self._kernel_cache = _initialize_caches(
                train_targets,
                train_noise_term.transpose(-1, -2),
                _stencils(initial_kxx),
                self.covar_module.num_inducing,
                create_w_cache=True,
            )
        else:
            self._kernel_cache = kernel_cache

    # ------------------------------------------------------------------ helpers
    @property
    def _dtype(self):
        return self._kernel_cache["interpolation_cache"].dtype

    @property
    def _num_models(self):
        return self._kernel_cache["interpolation_cache"].shape[0]

    def mean_module(self, X):
        return torch.zeros(X.shape[:-1], dtype=X.dtype, device=X.device)

    def _second_noise(self, o=None):
        noise = self.likelihood.second_noise_covar.noise.to(self._dtype)
        if o is None:
            return noise
        flat = noise.reshape(-1)
        return flat[o if flat.numel() > 1 else 0]

    def _root_panels(self):
        wtw = self._kernel_cache["WtW"]
        wtw.root_decomposition()
        return wtw._panels(wtw.root)

    def _update_cache_dicts(self, targets, noise_diagonal, stencils, inplace=False):
        """``_update_cache_dicts`` (``:155-171``): targets [q,t], noise_diagonal [t,q]."""
        idx, val = stencils
        m = self.covar_module.num_inducing
        new = _initialize_caches(targets, noise_diagonal, stencils, m, create_w_cache=False)
        old = self._kernel_cache
        updated = {}
        for key in old.keys():
            if key != "WtW":
                updated[key] = old[key].add_(new[key]) if inplace else new[key] + old[key]
            else:
                # every cache but "WtW" is additive; the root decomposition gets the rank-q update
                nd = noise_diagonal.expand(old["interpolation_cache"].shape[0], -1) if noise_diagonal.dim() == 2 \
                    else noise_diagonal
                root_noise = nd.clamp_min(1e-7) ** 0.5                                       # :163
                new_w_dinv = val.unsqueeze(0) / root_noise.unsqueeze(-1)                    # [t,q,s]   :168
                updated[key] = old[key].update_sparse(idx, new_w_dinv.contiguous(), inplace=inplace)     # :169
        return updated

    # ------------------------------------------------------------------ forward
    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def forward(self, X, **kwargs):
        if self.training:
            # a dummy: the real action happens in the MLL (``:174-203``)
            if X is not None:
                mean = self.mean_module(X)
                covar = self.covar_module(X)
            else:
                batch_shape = torch.Size((self._batch_shape,)) if type(self._batch_shape) is not torch.Size \
                    else self._batch_shape
                mean_shape = batch_shape + torch.Size((self.num_data,))
                dev = self._kernel_cache["interpolation_cache"].device
                mean = torch.zeros(*mean_shape, dtype=self._dtype, device=dev)
                covar = ZeroLazyTensor(*mean_shape, self.num_data, dtype=self._dtype, device=dev)
            if (
                mean.dim() < covar.dim()
                and (self._batch_shape != torch.Size() and mean.shape != covar.shape[:-1])
            ):
                mean = mean.unsqueeze(0).repeat(covar.shape[0], *[1] * (covar.dim() - 1))
            return MultivariateNormal(mean, covar)

        lazy_kernel = self.covar_module(X).evaluate_kernel()
        idx, val = lazy_kernel.left_interp_indices, lazy_kernel.left_interp_values
        xb = idx.shape[:-2]
        idx2, val2 = idx.reshape(-1, idx.shape[-1]), val.reshape(-1, val.shape[-1])
        cache = self.prediction_cache
        t = self._num_models
        # W* (K b - K L Q^-1 c)  (:206-210).  For few test points the m-vector pred_mean is never materialised:
        # W* K b - (W* K L) (Q^-1 c) only gathers the stencil rows of K L (shared with the covariance below).
        gather_form = (not cache.has_pred_mean()) and idx2.numel() * 2 < self.covar_module.num_inducing
        Ts = [None] * t
        if gather_form:
            means = []
            for o in range(t):
                Ts[o] = ops.left_interp(idx2, val2, cache["KL"][o]).transpose(-1, -2)          # r x q*
                means.append(ops.left_interp(idx2, val2, self.Kuu_response[o]) - Ts[o].transpose(-1, -2) @ cache["qmat_solve"][o])
            pred_mean = torch.stack(means)
        else:
            pred_mean = torch.stack([ops.left_interp(idx2, val2, cache["pred_mean"][o]) for o in range(t)])
        pred_mean = pred_mean.reshape(t, *xb, idx.shape[-2], 1)

        if skip_posterior_variances.off():
            if "pred_cov" not in cache:
                cache["pred_cov"] = self._make_predictive_covar()
            inner = cache["pred_cov"]
            if fast_pred_samples.off():
                covs = []
                for o in range(t):
                    scale = self._second_noise(o) if self.has_learnable_noise else None                   # :227-228
                    if len(xb) == 0:
                        covs.append(PredictiveCovar(idx2, val2, inner[o], scale, T=Ts[o]))
                    else:
                        n = idx.shape[-2]
                        covs.append(BatchLazyTensor([PredictiveCovar(idx2[b * n:(b + 1) * n], val2[b * n:(b + 1) * n],
                                                                     inner[o], scale) for b in range(xb.numel())]))
                pred_cov = BatchLazyTensor(covs)
            else:
                # fast_pred_samples (:229-243): the reference samples from a Lanczos root of the m x m grid-space
                # covariance truncated to q* columns (random start vector, approximate).  The q* x q* test covariance
                # is formed exactly here anyway, so the root handed to the sampler is its exact Cholesky factor.
                from ..lazy.lazy_tensor import psd_safe_cholesky
                covs = []
                n = idx.shape[-2]
                for o in range(t):
                    scale = self._second_noise(o) if self.has_learnable_noise else None
                    blocks = [PredictiveCovar(idx2[b * n:(b + 1) * n], val2[b * n:(b + 1) * n], inner[o], scale,
                                              T=Ts[o] if len(xb) == 0 else None) for b in range(max(1, xb.numel()))]
                    roots = [RootLazyTensor(psd_safe_cholesky(blk.evaluate())) for blk in blocks]
                    covs.append(roots[0] if len(xb) == 0 else BatchLazyTensor(roots))
                pred_cov = BatchLazyTensor(covs)
        else:
            pred_cov = ZeroLazyTensor(*lazy_kernel.shape, dtype=val.dtype, device=val.device)

        pred_mean = pred_mean[..., 0]
        if len(xb) > 0 and t == 1:
            pred_mean = pred_mean[0]
            if isinstance(pred_cov, BatchLazyTensor):
                pred_cov = pred_cov[0]
        elif self._batch_shape == torch.Size() and X.dim() == 2:
            pred_mean = pred_mean[0]
            if isinstance(pred_cov, BatchLazyTensor):
                pred_cov = pred_cov[0]
        return MultivariateNormal(pred_mean, pred_cov)

    # ------------------------------------------------------------------ conditioning
    def prestart_condition(self, X, Y, noise=None):
        """settings.overlap_root_update: start the inverse-root half of the in-place conditioning on (X, Y) on a side
        stream (see ``UpdatedRootLazyTensor.prestart_update_sparse``).  Must be followed by
        ``condition_on_observations(X, Y, noise, inplace=True)`` with the same arguments."""
        if noise is None:
            noise = torch.ones_like(Y)
        idx, val = _stencils(self.covar_module(X).evaluate_kernel())
        if (noise.shape[:-2] != self._batch_shape or noise.shape[-1] == 1) and noise.dim() < 3:
            nd = noise.transpose(-1, -2)
        else:
            nd = noise
        old = self._kernel_cache
        if nd.dim() == 2:
            nd = nd.expand(old["interpolation_cache"].shape[0], -1)
        new_w_dinv = val.unsqueeze(0) / (nd.clamp_min(1e-7) ** 0.5).unsqueeze(-1)          # as in _update_cache_dicts (:163-168)
        return old["WtW"].prestart_update_sparse(idx, new_w_dinv.contiguous())

    def condition_on_observations(self, X, Y, noise=None, inplace=False):
        if noise is None:
            noise = torch.ones_like(Y)
        lazy_kernel = self.covar_module(X).evaluate_kernel()
        if (noise.shape[:-2] != self._batch_shape or noise.shape[-1] == 1) and noise.dim() < 3:
            noise_for_update = noise.transpose(-1, -2)
        else:
            noise_for_update = noise
        new_kernel_cache = self._update_cache_dicts(Y, noise_for_update, _stencils(lazy_kernel), inplace=inplace)

        if inplace:
            self.num_data = self.num_data + X.shape[-2]
            if self._num_data_t is not None:
                self._num_data_t.add_(X.shape[-2])
            self._kernel_cache = new_kernel_cache
            self._dump_caches()
        else:
            new_gp = type(self)(
                covar_module=self.covar_module,
                kernel_cache=new_kernel_cache,
                learn_additional_noise=self.has_learnable_noise,
                likelihood=self.likelihood,
                num_data=self.num_data + X.shape[-2],
            )
            if self.training is False:
                new_gp.eval()
            return new_gp

    def get_fantasy_model(self, inputs, targets, noise_term, **kwargs):
        target_batch_shape = targets.shape[:-1]
        input_batch_shape = inputs.shape[:-2]
        tbdim, ibdim = len(target_batch_shape), len(input_batch_shape)
        if not (tbdim == ibdim + 1 or tbdim == ibdim):
            raise RuntimeError(
                f"Unsupported batch shapes: The target batch shape ({target_batch_shape}) must have either the "
                f"same dimension as or one more dimension than the input batch shape ({input_batch_shape})"
            )
        if ibdim > 1:
            raise RuntimeError(f"Unsupported batch shapes: fantasy inputs may carry one batch dimension, got {input_batch_shape}")
        if ibdim == 1:
            # candidate sets that differ per batch element: predictive-space fantasy (fantasy.py) instead of the
            # reference's per-element copies of WtW and both root panels (``_expand_batch``, :139-159)
            from .fantasy import PredictiveSpaceFantasy
            Y = targets.unsqueeze(-1)                                            # [(nf,) b, q, 1]  (:328)
            noise = noise_term if noise_term.dim() == Y.dim() else noise_term.unsqueeze(-1)
            if noise.dim() == 4:
                noise = noise[0]
            return PredictiveSpaceFantasy(self, inputs, Y, noise)
        return self.condition_on_observations(inputs, targets.reshape(inputs.shape[-2], -1), noise_term, inplace=False)

    # ------------------------------------------------------------------ cached properties (``:334-383``)
    @property
    @cached(name="Kuu")
    def Kuu(self):
        Kuu = self.covar_module._inducing_forward(last_dim_is_batch=False)
        t = self._num_models
        items = list(Kuu.items) if isinstance(Kuu, BatchLazyTensor) else [Kuu] * t
        if len(items) != t:
            raise RuntimeError(f"kernel batch ({len(items)}) does not match the number of outputs ({t})")
        out = []
        for o, K in enumerate(items):
            K = KroneckerToeplitzLazyTensor(K.cols.to(self._dtype), K.sizes,
                                            None if K.dirs is None else K.dirs.to(self._dtype))
            if self.has_learnable_noise:
                # learnable noise: K_uu / sigma^2, so that Q = I + L^T (K_uu / sigma^2) L (``:338-340``)
                K = K / self._second_noise(o)
            out.append(K)
        return BatchLazyTensor(out)

    @property
    @cached(name="current_inducing_compression_matrix")
    def current_inducing_compression_matrix(self):
        Ls = self._root_panels()
        return BatchLazyTensor([K @ PanelLazyTensor(L) for K, L in zip(self.Kuu.items, Ls)])

    @property
    @cached(name="current_qmatrix")
    def current_qmatrix(self):
        Ls = self._root_panels()
        KLs = self.current_inducing_compression_matrix.items
        return BatchLazyTensor([(PanelLazyTensor(L).transpose(-1, -2) @ KL).add_jitter(1.0) for L, KL in zip(Ls, KLs)])

    @property
    @cached(name="root_space_projection")
    def root_space_projection(self):
        Ls = self._root_panels()
        return torch.stack([PanelLazyTensor(L).transpose(-1, -2) @ Kb for L, Kb in zip(Ls, self.Kuu_response)])

    @property
    @cached(name="Kuu_response")
    def Kuu_response(self):
        return self.Kuu.matmul(self._kernel_cache["interpolation_cache"])

    @property
    @cached(name="prediction_cache")
    def prediction_cache(self):
        KLs = [kl.evaluate() for kl in self.current_inducing_compression_matrix.items]
        if overlap_root_update.on() and ops.overlap_capable(KLs[0]) and max(KL.shape[-1] for KL in KLs) <= max_cholesky_size.value():
            # c = L^T (K b) is one HBM-bound pass over L (and its backward another one): both go to a side stream, under the
            # tensor-bound Gram L^T (K L) / its backward panel GEMM on this one; joined before the solve needs c
            self.Kuu_response              # (its small Kronecker passes stay on this stream)
            with ops.side_section(KLs[0].device):
                self.root_space_projection
            for Q in self.current_qmatrix.items:
                Q.cholesky()
            ops.join_side(KLs[0].device)
        qmat_solve = self.current_qmatrix.inv_matmul(self.root_space_projection)
        prediction_cache = _PredictionCache(self.Kuu_response, KLs, qmat_solve)
        if skip_posterior_variances.off():
            prediction_cache["pred_cov"] = self._make_predictive_covar(self.current_qmatrix, self.Kuu, KLs)
        return prediction_cache

    def _make_predictive_covar(self, qmatrix=None, Kuu=None, Kuu_Lmat=None):
        if qmatrix is None:
            qmatrix = self.current_qmatrix
        if Kuu is None:
            Kuu = self.Kuu
        if Kuu_Lmat is None:
            Kuu_Lmat = [kl.evaluate() for kl in self.current_inducing_compression_matrix.items]
        # fast_pred_var on/off only changes how the same operator is *represented* in the reference (root form
        # K L Q^-1/2 vs a dense m x m product); both are the operator below.
        return BatchLazyTensor([WoodburyInnerCovar(K, KL, Q) for K, KL, Q in zip(Kuu.items, Kuu_Lmat, qmatrix.items)])

    def _dump_caches(self):
        fixed_cache_names = ["current_qmatrix", "current_inducing_compression_matrix", "prediction_cache",
                             "root_space_projection", "Kuu_response", "Kuu"]
        for name in fixed_cache_names:
            try:
                pop_from_cache(self, name)
            except CachingError:
                pass

    def zero_grad(self, set_to_none=True):
        self._dump_caches()
        return super().zero_grad(set_to_none=set_to_none)

    def set_train_data(self, train_inputs, train_targets, train_noise_term):
        initial_kxx = self.covar_module(train_inputs).evaluate_kernel()
        self._kernel_cache = _initialize_caches(
            train_targets,
            train_noise_term.transpose(-1, -2),
            _stencils(initial_kxx),
            self.covar_module.num_inducing,
            create_w_cache=True,
        )
        self.num_data = train_inputs.shape[-2]
        if self._num_data_t is not None:
            self._num_data_t.fill_(float(self.num_data))

    # ------------------------------------------------------------------ state (de)serialisation
    # The reference keeps the WISKI caches outside ``state_dict`` (a reloaded model silently loses every
    # observation).  Here they travel as the module's extra state, so ``model.state_dict()`` /
    # ``load_state_dict()`` round-trip the full posterior: caches, root / inverse-root panels and the counter.
    def get_extra_state(self):
        wtw = self._kernel_cache["WtW"]
        cpu = lambda t: None if t is None else t.detach().to("cpu", copy=True)     # a snapshot, never a view
        return {
            "num_data": int(self.num_data),
            "response_cache": cpu(self._kernel_cache["response_cache"]),
            "interpolation_cache": cpu(self._kernel_cache["interpolation_cache"]),
            "D_logdet": cpu(self._kernel_cache["D_logdet"]),
            "WtW": {"tensor": cpu(wtw.tensor), "root": cpu(wtw.root), "inv_root": cpu(wtw.inv_root)},
        }

    def set_extra_state(self, state):
        dev = self._kernel_cache["interpolation_cache"].device
        mv = lambda t: None if t is None else t.to(dev, copy=True)
        self._kernel_cache = {
            "response_cache": mv(state["response_cache"]),
            "interpolation_cache": mv(state["interpolation_cache"]),
            "WtW": UpdatedRootLazyTensor(mv(state["WtW"]["tensor"]), initial_is_root=False, root=mv(state["WtW"]["root"]),
                                         inv_root=mv(state["WtW"]["inv_root"])),
            "D_logdet": mv(state["D_logdet"]),
        }
        self.num_data = int(state["num_data"])
        if self._num_data_t is not None:
            self._num_data_t.fill_(float(self.num_data))
        self._dump_caches()

    def to(self, *args, **kwargs):
        device = args[0] if args else kwargs.get("device")
        if torch.is_tensor(device):
            device = device.device
        if self._kernel_cache is not None and device is not None:
            for key in self._kernel_cache:
                self._kernel_cache[key] = self._kernel_cache[key].to(device)
        return super().to(*args, **kwargs)
