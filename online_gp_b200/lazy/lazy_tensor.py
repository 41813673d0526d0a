"""Minimal LazyTensor operator algebra backed by the CUDA kernels.

GPyTorch is not a dependency of this package (it cannot be installed in the build environment), so
the handful of ``gpytorch.lazy`` classes the WISKI path touches are provided here with the same protocol
(``_size / _matmul / _transpose_nonbatch / _solve / evaluate``) and the same derived surface the reference calls
(``matmul, @, transpose, add_jitter, inv_matmul, inv_quad_logdet, root_decomposition, root_inv_decomposition,
evaluate, detach, diag``; ``online_gp/models/batched_fixed_noise_online_gp.py:346-404``,
``online_gp/mlls/batched_woodbury_marginal_log_likelihood.py:27-30``).  Dispatch rules follow SURVEY.md App. A.5:
size <= ``max_cholesky_size`` -> dense Cholesky (``psd_safe_cholesky`` jitter escalation), otherwise ``_solve`` (CG).

Operators here are 2-D (one GP output); the model loops over outputs and stacks (``BatchLazyTensor``).
"""
import math
import warnings

import torch

from .. import ops, settings


class NumericalWarning(RuntimeWarning):
    pass


class NotPSDError(RuntimeError):
    pass


class NanError(RuntimeError):
    pass


def psd_safe_cholesky(A, jitter=None, max_tries=3):
    """GPyTorch ``psd_safe_cholesky`` (App. A.5): plain Cholesky, then jitter * 10**i on the diagonal, i < max_tries."""
    L, info = torch.linalg.cholesky_ex(A)
    if not bool(info.any()):
        return L
    if bool(torch.isnan(A).any()):
        raise NanError(f"cholesky: {int(torch.isnan(A).sum())} of {A.numel()} elements of the matrix are NaN.")
    if jitter is None:
        jitter = settings.cholesky_jitter.value(A.dtype)
    Ap, prev = A.clone(), 0.0
    for i in range(max_tries):
        jn = jitter * (10 ** i)
        Ap.diagonal(dim1=-2, dim2=-1).add_(jn - prev)
        prev = jn
        L, info = torch.linalg.cholesky_ex(Ap)
        if not bool(info.any()):
            warnings.warn(f"A not p.d., added jitter of {jn:.1e} to the diagonal", NumericalWarning)
            return L
    raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to {jn:.1e}.")


class _InvQuadFn(torch.autograd.Function):
    """b^T A^-1 b per column of b, given a solve x = A^-1 b computed elsewhere (CG): the forward value uses x, the
    backward is  dA = -x g x^T,  db = 2 x g  (GPyTorch ``InvQuadLogDet.backward`` without the probe vectors)."""

    @staticmethod
    def forward(ctx, A, b, x):
        ctx.save_for_backward(x)
        return (x * b).sum(-2)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        xg = x * g.unsqueeze(-2)
        gA = -(xg @ x.transpose(-1, -2)) if ctx.needs_input_grad[0] else None
        gb = 2.0 * xg if ctx.needs_input_grad[1] else None
        return gA, gb, None


class LazyTensor:
    # ---- protocol (override)
    def _size(self):
        raise NotImplementedError

    def _matmul(self, rhs):
        raise NotImplementedError

    def _transpose_nonbatch(self):
        raise NotImplementedError

    def _solve(self, rhs, preconditioner=None, num_tridiag=0):
        """Iterative solve used beyond the Cholesky regime; dense operators fall back to their Cholesky factor."""
        return torch.cholesky_solve(rhs, self.cholesky())

    def evaluate(self):
        n = self.shape[-1]
        return self._matmul(torch.eye(n, dtype=self.dtype, device=self.device))

    @property
    def dtype(self):
        raise NotImplementedError

    @property
    def device(self):
        raise NotImplementedError

    # ---- derived surface
    @property
    def shape(self):
        return self._size()

    def size(self, dim=None):
        s = self._size()
        return s if dim is None else s[dim]

    def dim(self):
        return len(self._size())

    ndimension = dim

    @property
    def batch_shape(self):
        return self._size()[:-2]

    @property
    def matrix_shape(self):
        return self._size()[-2:]

    @property
    def is_square(self):
        return self.shape[-1] == self.shape[-2]

    def transpose(self, d1, d2):
        nd = self.dim()
        d1, d2 = d1 % nd, d2 % nd
        if {d1, d2} != {nd - 1, nd - 2}:
            raise NotImplementedError("only the non-batch transpose is supported")
        return self._transpose_nonbatch()

    def t(self):
        return self._transpose_nonbatch()

    def matmul(self, other):
        if isinstance(other, LazyTensor):
            return MatmulLazyTensor(self, other)
        if other.dim() == 1:
            return self._matmul(other.unsqueeze(-1)).squeeze(-1)
        return self._matmul(other)

    __matmul__ = matmul

    def __rmatmul__(self, other):
        # tensor @ lazy  ==  (lazy^T @ tensor^T)^T
        return self._transpose_nonbatch().matmul(other.transpose(-1, -2)).transpose(-1, -2)

    def add_jitter(self, jitter_val=1e-3):
        return AddedDiagLazyTensor(self, jitter_val)

    def __add__(self, other):
        return SumLazyTensor(self, lazify(other))

    def __sub__(self, other):
        return SumLazyTensor(self, lazify(other), sign=-1.0)

    def __mul__(self, other):
        return ConstantMulLazyTensor(self, other)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return ConstantMulLazyTensor(self, 1.0 / other)

    def detach(self):
        return self

    def diag(self):
        return self.evaluate().diagonal(dim1=-2, dim2=-1)

    def cholesky(self, upper=False):
        if not hasattr(self, "_chol_cache"):
            self._chol_cache = psd_safe_cholesky(self.evaluate())
        return self._chol_cache.transpose(-1, -2) if upper else self._chol_cache

    def _use_cholesky(self):
        return self.shape[-1] <= settings.max_cholesky_size.value()

    def inv_matmul(self, right_tensor, left_tensor=None):
        squeeze = right_tensor.dim() == 1
        rhs = right_tensor.unsqueeze(-1) if squeeze else right_tensor
        if self._use_cholesky():
            sol = torch.cholesky_solve(rhs, self.cholesky())
        else:
            sol = self._solve(rhs)
        if squeeze:
            sol = sol.squeeze(-1)
        return sol if left_tensor is None else left_tensor @ sol

    def inv_quad_logdet(self, inv_quad_rhs=None, logdet=False, reduce_inv_quad=True):
        """(rhs^T A^-1 rhs, log|A|) on the Cholesky path; ``skip_logdet_forward`` zeroes the logdet *value* only."""
        inv_quad_term = logdet_term = None
        if not self._use_cholesky():
            return self._inv_quad_logdet_iterative(inv_quad_rhs, logdet, reduce_inv_quad)
        Lc = self.cholesky()
        if inv_quad_rhs is not None:
            rhs = inv_quad_rhs.unsqueeze(-1) if inv_quad_rhs.dim() == 1 else inv_quad_rhs
            half = torch.linalg.solve_triangular(Lc, rhs, upper=False)
            inv_quad_term = (half * half).sum(-2)
            if reduce_inv_quad:
                inv_quad_term = inv_quad_term.sum(-1)
        if logdet:
            logdet_term = 2.0 * Lc.diagonal(dim1=-2, dim2=-1).log().sum(-1)
            if settings.skip_logdet_forward.on():
                logdet_term = logdet_term - logdet_term.detach()
        return inv_quad_term, logdet_term

    def _inv_quad_logdet_iterative(self, inv_quad_rhs, logdet, reduce_inv_quad):
        """Beyond ``max_cholesky_size`` (GPyTorch: mBCG solves + stochastic Lanczos quadrature, App. A.5).  The solve is
        the operator's CG driver, differentiable through  d(b^T A^-1 b) = -x^T dA x + 2 x^T db  with x = A^-1 b
        (``_InvQuadFn``).  log|A|: GPyTorch estimates it with ``num_trace_samples`` random probes (non-deterministic, a
        few per cent off); the operators of this path are at most max_root_decomposition_size wide, so it is taken
        exactly from one dense factorisation instead — deterministic, and differentiable the ordinary way."""
        inv_quad_term = logdet_term = None
        dense = self.evaluate() if (logdet or (inv_quad_rhs is not None and torch.is_grad_enabled())) else None
        if inv_quad_rhs is not None:
            rhs = inv_quad_rhs.unsqueeze(-1) if inv_quad_rhs.dim() == 1 else inv_quad_rhs
            sol = self._solve(rhs.detach())
            if dense is not None and (dense.requires_grad or rhs.requires_grad):
                inv_quad_term = _InvQuadFn.apply(dense, rhs, sol.detach())
            else:
                inv_quad_term = (sol * rhs).sum(-2)
            if reduce_inv_quad:
                inv_quad_term = inv_quad_term.sum(-1)
        if logdet:
            Lc = psd_safe_cholesky(dense)
            logdet_term = 2.0 * Lc.diagonal(dim1=-2, dim2=-1).log().sum(-1)
            if settings.skip_logdet_forward.on():
                logdet_term = logdet_term - logdet_term.detach()
        return inv_quad_term, logdet_term

    def root_decomposition(self, method=None):
        return RootLazyTensor(self.cholesky())

    def root_inv_decomposition(self, initial_vectors=None, test_vectors=None):
        Lc = self.cholesky()
        eye = torch.eye(Lc.shape[-1], dtype=Lc.dtype, device=Lc.device)
        return RootLazyTensor(torch.linalg.solve_triangular(Lc, eye, upper=False).transpose(-1, -2))

    def to(self, device):
        return self


class NonLazyTensor(LazyTensor):
    """Dense matrix (small: r x r, q x q ...) — products go through torch.matmul."""

    def __init__(self, tensor):
        self.tensor = tensor

    def _size(self):
        return self.tensor.shape

    def _matmul(self, rhs):
        return self.tensor @ rhs

    def _transpose_nonbatch(self):
        return NonLazyTensor(self.tensor.transpose(-1, -2))

    def evaluate(self):
        return self.tensor

    def detach(self):
        return NonLazyTensor(self.tensor.detach())

    def to(self, device):
        return NonLazyTensor(self.tensor.to(device))

    dtype = property(lambda self: self.tensor.dtype)
    device = property(lambda self: self.tensor.device)


class PanelLazyTensor(LazyTensor):
    """m x r panel (root L, inverse root B, K L ...); products run in the panel kernels."""

    def __init__(self, panel):
        self.panel = panel

    def _size(self):
        return self.panel.shape

    def _matmul(self, rhs):
        return ops.panel_rmul(self.panel, rhs)

    def _transpose_nonbatch(self):
        return PanelTLazyTensor(self.panel)

    def evaluate(self):
        return self.panel

    def detach(self):
        return PanelLazyTensor(self.panel.detach())

    def to(self, device):
        return PanelLazyTensor(self.panel.to(device))

    dtype = property(lambda self: self.panel.dtype)
    device = property(lambda self: self.panel.device)


class PanelTLazyTensor(LazyTensor):
    """Transposed panel (r x m): ``P^T @ X`` is the Gram kernel."""

    def __init__(self, panel):
        self.panel = panel

    def _size(self):
        return torch.Size((self.panel.shape[1], self.panel.shape[0]))

    def _matmul(self, rhs):
        return ops.gram(self.panel, rhs)

    def _transpose_nonbatch(self):
        return PanelLazyTensor(self.panel)

    def evaluate(self):
        return self.panel.transpose(-1, -2)

    def detach(self):
        return PanelTLazyTensor(self.panel.detach())

    dtype = property(lambda self: self.panel.dtype)
    device = property(lambda self: self.panel.device)


def lazify(obj):
    if isinstance(obj, LazyTensor):
        return obj
    if torch.is_tensor(obj):
        return NonLazyTensor(obj)
    raise TypeError(f"cannot lazify {type(obj)}")


def delazify(obj):
    return obj.evaluate() if isinstance(obj, LazyTensor) else obj


class RootLazyTensor(LazyTensor):
    """R R^T with ``.root`` a LazyTensor (panel when tall)."""

    def __init__(self, root):
        if torch.is_tensor(root):
            root = PanelLazyTensor(root) if root.shape[-2] > root.shape[-1] or root.shape[-2] >= 64 else NonLazyTensor(root)
        self.root = root

    def _size(self):
        n = self.root.shape[-2]
        return torch.Size((*self.root.shape[:-2], n, n))

    def _matmul(self, rhs):
        return self.root._matmul(self.root._transpose_nonbatch()._matmul(rhs))

    def _transpose_nonbatch(self):
        return self

    def evaluate(self):
        R = self.root.evaluate()
        return R @ R.transpose(-1, -2)

    def diag(self):
        R = self.root.evaluate()
        return (R * R).sum(-1)

    def root_decomposition(self, method=None):
        return self

    dtype = property(lambda self: self.root.dtype)
    device = property(lambda self: self.root.device)


class DiagLazyTensor(LazyTensor):
    def __init__(self, diag):
        self._diag = diag

    def _size(self):
        n = self._diag.shape[-1]
        return torch.Size((*self._diag.shape[:-1], n, n))

    def _matmul(self, rhs):
        return self._diag.unsqueeze(-1) * rhs

    def _transpose_nonbatch(self):
        return self

    def evaluate(self):
        return torch.diag_embed(self._diag)

    def diag(self):
        return self._diag

    def inv_matmul(self, right_tensor, left_tensor=None):
        res = right_tensor / (self._diag.unsqueeze(-1) if right_tensor.dim() > self._diag.dim() else self._diag)
        return res if left_tensor is None else left_tensor @ res

    def logdet(self):
        return self._diag.log().sum(-1)

    dtype = property(lambda self: self._diag.dtype)
    device = property(lambda self: self._diag.device)


class ZeroLazyTensor(LazyTensor):
    def __init__(self, *sizes, dtype=None, device=None):
        self._sizes = torch.Size(sizes)
        self._dtype = dtype or torch.get_default_dtype()
        self._device = device or torch.device("cpu")

    def _size(self):
        return self._sizes

    def _matmul(self, rhs):
        return torch.zeros(*self._sizes[:-1], rhs.shape[-1], dtype=rhs.dtype, device=rhs.device)

    def _transpose_nonbatch(self):
        return ZeroLazyTensor(*self._sizes[:-2], self._sizes[-1], self._sizes[-2], dtype=self._dtype, device=self._device)

    def evaluate(self):
        return torch.zeros(*self._sizes, dtype=self._dtype, device=self._device)

    def diag(self):
        return torch.zeros(*self._sizes[:-1], dtype=self._dtype, device=self._device)

    dtype = property(lambda self: self._dtype)
    device = property(lambda self: self._device)


class ConstantMulLazyTensor(LazyTensor):
    def __init__(self, base, constant):
        self.base, self.constant = base, constant

    def _size(self):
        return self.base._size()

    def _matmul(self, rhs):
        return self.base._matmul(rhs) * self.constant

    def _transpose_nonbatch(self):
        return ConstantMulLazyTensor(self.base._transpose_nonbatch(), self.constant)

    def evaluate(self):
        return self.base.evaluate() * self.constant

    def diag(self):
        return self.base.diag() * self.constant

    def detach(self):
        c = self.constant.detach() if torch.is_tensor(self.constant) else self.constant
        return ConstantMulLazyTensor(self.base.detach(), c)

    dtype = property(lambda self: self.base.dtype)
    device = property(lambda self: self.base.device)


class SumLazyTensor(LazyTensor):
    def __init__(self, a, b, sign=1.0):
        self.a, self.b, self.sign = a, b, sign

    def _size(self):
        return self.a._size()

    def _matmul(self, rhs):
        return self.a._matmul(rhs) + self.sign * self.b._matmul(rhs)

    def _transpose_nonbatch(self):
        return SumLazyTensor(self.a._transpose_nonbatch(), self.b._transpose_nonbatch(), self.sign)

    def evaluate(self):
        return self.a.evaluate() + self.sign * self.b.evaluate()

    def diag(self):
        return self.a.diag() + self.sign * self.b.diag()

    def detach(self):
        return SumLazyTensor(self.a.detach(), self.b.detach(), self.sign)

    dtype = property(lambda self: self.a.dtype)
    device = property(lambda self: self.a.device)


class AddedDiagLazyTensor(LazyTensor):
    """A + jitter * I  (``add_jitter``)."""

    def __init__(self, base, jitter):
        self.base, self.jitter = base, jitter

    def _size(self):
        return self.base._size()

    def _matmul(self, rhs):
        return self.base._matmul(rhs) + self.jitter * rhs

    def _transpose_nonbatch(self):
        return AddedDiagLazyTensor(self.base._transpose_nonbatch(), self.jitter)

    def evaluate(self):
        E = self.base.evaluate()
        return E + self.jitter * torch.eye(E.shape[-1], dtype=E.dtype, device=E.device)

    def detach(self):
        return AddedDiagLazyTensor(self.base.detach(), self.jitter)

    dtype = property(lambda self: self.base.dtype)
    device = property(lambda self: self.base.device)


class MatmulLazyTensor(LazyTensor):
    """Lazy product; evaluation dispatches to the structured kernels of the left factor."""

    def __init__(self, left, right):
        self.left, self.right = lazify(left), lazify(right)
        self._eval = None

    def _size(self):
        return torch.Size((*self.left.shape[:-1], self.right.shape[-1]))

    def _matmul(self, rhs):
        return self.left._matmul(self.right._matmul(rhs))

    def _transpose_nonbatch(self):
        return MatmulLazyTensor(self.right._transpose_nonbatch(), self.left._transpose_nonbatch())

    def evaluate(self):
        if self._eval is None:
            self._eval = self.left._matmul(self.right.evaluate())
        return self._eval

    def add_jitter(self, jitter_val=1e-3):
        # L^T (K L) + jitter I : the WISKI Q matrix — keep the panels so the CG path can use the fused MVM
        if isinstance(self.left, PanelTLazyTensor):
            return PanelGramLazyTensor(self.left.panel, self.right.evaluate(), jitter_val)
        return AddedDiagLazyTensor(self, jitter_val)

    def detach(self):
        return MatmulLazyTensor(self.left.detach(), self.right.detach())

    dtype = property(lambda self: self.left.dtype)
    device = property(lambda self: self.left.device)


class PanelGramLazyTensor(LazyTensor):
    """Q = jitter * I + L^T KL for two m x r panels (``current_qmatrix``,
    ``online_gp/models/batched_fixed_noise_online_gp.py:350-355``).

    * r <= max_cholesky_size: Q is formed once by the Gram kernel and factorised (what GPyTorch dispatches to in
      every shipped config);
    * otherwise ``_matmul`` is the fused one-pass panel MVM and ``_solve`` the CG driver built on it (App. A.5).
    """

    def __init__(self, L, KL, jitter=1.0):
        self.L, self.KL, self.jitter = L, KL, jitter
        self._eval = None

    def _size(self):
        r = self.L.shape[1]
        return torch.Size((r, r))

    def _matmul(self, rhs):
        if self.jitter == 1.0 and self.L.shape[1] <= 1024 and not (torch.is_grad_enabled() and self.KL.requires_grad):
            return ops.q_matvec(self.L, self.KL.detach(), rhs)
        return ops.gram(self.L, ops.panel_rmul(self.KL, rhs)) + self.jitter * rhs

    def _transpose_nonbatch(self):
        return self     # symmetric (K symmetric)

    def evaluate(self):
        if self._eval is None:
            G = ops.gram(self.L, self.KL, symmetric=True)       # L^T (K L) with K symmetric
            self._eval = G + self.jitter * torch.eye(G.shape[-1], dtype=G.dtype, device=G.device)
        return self._eval

    def cholesky(self, upper=False):
        # Q = jitter I + L^T K L with K PSD: for jitter >= 1 it is positive definite by construction, so the
        # factorisation is issued without the host-side `info` check (one device->host sync less per step);
        # a breakdown (NaN inputs) still surfaces as NaNs downstream.
        if not hasattr(self, "_chol_cache"):
            if self.jitter >= 1.0:
                self._chol_cache, _ = torch.linalg.cholesky_ex(self.evaluate(), check_errors=False)
            else:
                self._chol_cache = psd_safe_cholesky(self.evaluate())
        return self._chol_cache.transpose(-1, -2) if upper else self._chol_cache

    def _solve(self, rhs, preconditioner=None, num_tridiag=0):
        if self.jitter != 1.0 or self.L.shape[1] > 1024:
            return torch.cholesky_solve(rhs, self.cholesky())
        tol = settings.eval_cg_tolerance.value() if not torch.is_grad_enabled() else settings.cg_tolerance.value()
        x, iters, resid = ops.cg_solve(self.L, self.KL.detach(), rhs.detach(), tol=tol,
                                       max_iter=settings.max_cg_iterations.value())
        if resid > tol:
            warnings.warn(f"CG terminated in {iters} iterations with average residual norm {resid:.3e} which is "
                          f"larger than the tolerance of {tol} specified by cg_tolerance.", NumericalWarning)
        self.last_cg = (iters, resid)
        return x

    def detach(self):
        return PanelGramLazyTensor(self.L.detach(), self.KL.detach(), self.jitter)

    dtype = property(lambda self: self.L.dtype)
    device = property(lambda self: self.L.device)


class KroneckerToeplitzLazyTensor(LazyTensor):
    """K_uu = kron_i Toeplitz(cols[i]) — stands in for ``KroneckerProductLazyTensor(ToeplitzLazyTensor...)``
    (SURVEY App. A.3/A.4); ``_matmul`` is the Kronecker-Toeplitz CUDA kernel, differentiable w.r.t. ``cols``."""

    def __init__(self, cols, sizes, dirs=None):
        # dirs (optional, [d,gmax], no grad): d cols[i] / d lengthscale_i — enables the directional backward pass
        self.cols, self.sizes, self.dirs = cols, tuple(int(s) for s in sizes), dirs
        self.m = 1
        for s in self.sizes:
            self.m *= s

    def _size(self):
        return torch.Size((self.m, self.m))

    def _matmul(self, rhs):
        return ops.kron_toeplitz_matmul(self.cols, self.sizes, rhs, dirs=self.dirs)

    def _transpose_nonbatch(self):
        return self

    def __truediv__(self, other):
        # fold the scalar into the first factor's column (Kuu / sigma^2, batched_fixed_noise_online_gp.py:340)
        scale = torch.ones_like(self.cols[:, :1])
        scale = torch.cat([1.0 / other.reshape(1, 1).to(self.cols), scale[1:]], dim=0) if self.cols.shape[0] > 1 \
            else 1.0 / other.reshape(1, 1).to(self.cols)
        return KroneckerToeplitzLazyTensor(self.cols * scale, self.sizes, self.dirs)

    def __mul__(self, other):
        return self.__truediv__(1.0 / torch.as_tensor(other, dtype=self.cols.dtype, device=self.cols.device))

    def diag(self):
        d0 = self.cols[:, 0].prod()
        return d0.expand(self.m)

    def evaluate(self):
        if self.m > 16384:
            raise RuntimeError(f"refusing to densify K_uu with m={self.m}; use matmul")
        return super().evaluate()

    def detach(self):
        return KroneckerToeplitzLazyTensor(self.cols.detach(), self.sizes, self.dirs)

    dtype = property(lambda self: self.cols.dtype)
    device = property(lambda self: self.cols.device)


class InterpolatedLazyTensor(LazyTensor):
    """W_l K W_r^T with sparse interpolation stencils (GPyTorch ``InterpolatedLazyTensor``); what
    ``covar_module(X).evaluate_kernel()`` returns (``batched_fixed_noise_online_gp.py:143,205,261``)."""

    def __init__(self, base_lazy_tensor, left_interp_indices, left_interp_values, right_interp_indices=None,
                 right_interp_values=None):
        # ``base_lazy_tensor`` may be given as a zero-argument callable; it is then built on first access
        self._base = base_lazy_tensor
        self.left_interp_indices, self.left_interp_values = left_interp_indices, left_interp_values
        self.right_interp_indices = left_interp_indices if right_interp_indices is None else right_interp_indices
        self.right_interp_values = left_interp_values if right_interp_values is None else right_interp_values

    @property
    def base_lazy_tensor(self):
        if callable(self._base) and not isinstance(self._base, LazyTensor):
            self._base = self._base()
        return self._base

    def _size(self):
        return torch.Size((*self.left_interp_indices.shape[:-1], self.right_interp_indices.shape[-2]))

    def _flat(self, t):
        return t.reshape(-1, t.shape[-1])

    def _matmul(self, rhs):
        m = self.base_lazy_tensor.shape[-1]
        li, lv = self._flat(self.left_interp_indices), self._flat(self.left_interp_values)
        ri, rv = self._flat(self.right_interp_indices), self._flat(self.right_interp_values)
        up = ops.left_t_interp(ri, rv, rhs, m) if not rv.requires_grad else _scatter_dense(ri, rv, rhs, m)
        return ops.left_interp(li, lv, self.base_lazy_tensor._matmul(up))

    def _transpose_nonbatch(self):
        return InterpolatedLazyTensor(self._base, self.right_interp_indices, self.right_interp_values,
                                      self.left_interp_indices, self.left_interp_values)

    def _sparse_left_interp_t(self, indices, values):
        """Dense-able W^T (m x q) as a torch sparse COO tensor — kept for API parity (``_get_wmat_from_kernel``,
        ``batched_fixed_noise_online_gp.py:22-28``); the hot path never calls it."""
        q, s = indices.shape[-2:]
        m = self.base_lazy_tensor.shape[-1]
        cols = torch.arange(q, device=indices.device).unsqueeze(-1).expand(q, s)
        return torch.sparse_coo_tensor(torch.stack([indices.reshape(-1), cols.reshape(-1)]), values.reshape(-1),
                                       (m, q))

    def diag(self):
        return self.evaluate().diagonal(dim1=-2, dim2=-1)

    dtype = property(lambda self: self.left_interp_values.dtype)
    device = property(lambda self: self.left_interp_values.device)


def _scatter_dense(idx, val, rhs, m):
    """Differentiable W^T rhs (used only when the interpolation values carry grad: stem training)."""
    out = torch.zeros(m, rhs.shape[-1], dtype=rhs.dtype, device=rhs.device)
    contrib = (val.unsqueeze(-1) * rhs.unsqueeze(1)).reshape(-1, rhs.shape[-1])
    return out.index_add(0, idx.reshape(-1), contrib)


class BatchLazyTensor(LazyTensor):
    """A stack of independent 2-D operators (one per GP output) presenting a batched shape."""

    def __init__(self, items):
        self.items = list(items)

    def _size(self):
        return torch.Size((len(self.items), *self.items[0].shape))

    def __getitem__(self, i):
        return self.items[i]

    def _matmul(self, rhs):
        return torch.stack([it._matmul(rhs[i] if rhs.dim() == 3 else rhs) for i, it in enumerate(self.items)])

    def _transpose_nonbatch(self):
        return BatchLazyTensor([it._transpose_nonbatch() for it in self.items])

    def evaluate(self):
        return torch.stack([it.evaluate() for it in self.items])

    def diag(self):
        return torch.stack([it.diag() for it in self.items])

    def add_jitter(self, jitter_val=1e-3):
        return BatchLazyTensor([it.add_jitter(jitter_val) for it in self.items])

    def inv_matmul(self, right_tensor, left_tensor=None):
        return torch.stack([it.inv_matmul(right_tensor[i], None if left_tensor is None else left_tensor[i])
                            for i, it in enumerate(self.items)])

    def inv_quad_logdet(self, inv_quad_rhs=None, logdet=False, reduce_inv_quad=True):
        res = [it.inv_quad_logdet(None if inv_quad_rhs is None else inv_quad_rhs[i], logdet, reduce_inv_quad)
               for i, it in enumerate(self.items)]
        iq = None if res[0][0] is None else torch.stack([r[0] for r in res])
        ld = None if res[0][1] is None else torch.stack([r[1] for r in res])
        return iq, ld

    def detach(self):
        return BatchLazyTensor([it.detach() for it in self.items])

    dtype = property(lambda self: self.items[0].dtype)
    device = property(lambda self: self.items[0].device)
