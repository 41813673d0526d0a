"""Kernel modules the WISKI path is built on: constraints, RBF / Matern / Scale base kernels and the
GridInterpolationKernel (SKI) wrapper.

Stand-ins for ``gpytorch.kernels.{RBFKernel, MaternKernel, ScaleKernel, GridInterpolationKernel}`` and
``gpytorch.constraints`` with the same constructor arguments, parameter names (``raw_lengthscale``,
``raw_outputscale``), shapes and default initialisation (raw = 0), following SURVEY.md App. A.2/A.3 — used at
``online_gp/models/batched_fixed_noise_online_gp.py:107-120`` and ``experiments/bayesopt/bayesopt.py:65-78``.
Only what the SKI path needs is implemented: the per-dimension first Toeplitz column on the inducing grid
(``base_kernel(first_grid_point, grid, last_dim_is_batch=True)``, A.3) and the interpolation stencils.
"""
import math

import torch
from torch import nn
import torch.nn.functional as F

from . import ops, settings
from .lazy.lazy_tensor import InterpolatedLazyTensor, KroneckerToeplitzLazyTensor, LazyTensor


# ------------------------------------------------------------------ constraints
class Interval(nn.Module):
    def __init__(self, lower_bound, upper_bound):
        super().__init__()
        self.lower_bound, self.upper_bound = float(lower_bound), float(upper_bound)

    def transform(self, raw):
        return torch.sigmoid(raw) * (self.upper_bound - self.lower_bound) + self.lower_bound

    def inverse_transform(self, value):
        v = (value - self.lower_bound) / (self.upper_bound - self.lower_bound)
        return torch.log(v) - torch.log1p(-v)


class GreaterThan(Interval):
    def __init__(self, lower_bound):
        super().__init__(lower_bound, math.inf)

    def transform(self, raw):
        return F.softplus(raw) + self.lower_bound

    def inverse_transform(self, value):
        v = value - self.lower_bound
        return v + torch.log(-torch.expm1(-v))


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)


class GammaPrior(torch.distributions.Gamma):
    """``gpytorch.priors.GammaPrior(concentration, rate)`` (bayesopt.py:68,72)."""

    def log_prob(self, x):
        return super().log_prob(x)


class _PriorMixin:
    def named_priors(self):
        """Yield (name, prior, closure, setting_closure) like ``gpytorch.Module.named_priors``."""
        for mod_name, mod in self.named_modules():
            for pname, (prior, closure) in getattr(mod, "_priors", {}).items():
                yield (f"{mod_name}.{pname}" if mod_name else pname), prior, (lambda m=mod, c=closure: c(m)), None


# ------------------------------------------------------------------ base kernels
class Kernel(nn.Module, _PriorMixin):
    has_lengthscale = True

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), lengthscale_prior=None,
                 lengthscale_constraint=None, active_dims=None, **kwargs):
        super().__init__()
        self.ard_num_dims = ard_num_dims
        self.batch_shape = torch.Size(batch_shape) if not isinstance(batch_shape, int) else torch.Size([batch_shape])
        self._priors = {}
        if self.has_lengthscale:
            nd = 1 if ard_num_dims is None else ard_num_dims
            self.raw_lengthscale = nn.Parameter(torch.zeros(tuple(self.batch_shape) + (1, nd)))
            self.raw_lengthscale_constraint = lengthscale_constraint if lengthscale_constraint is not None else Positive()
            if lengthscale_prior is not None:
                self._priors["lengthscale_prior"] = (lengthscale_prior, lambda m: m.lengthscale)

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, value):
        value = torch.as_tensor(value, dtype=self.raw_lengthscale.dtype, device=self.raw_lengthscale.device)
        with torch.no_grad():
            self.raw_lengthscale.copy_(self.raw_lengthscale_constraint.inverse_transform(value).expand_as(self.raw_lengthscale))

    def base_1d(self, r):
        """Stationary profile as a function of |delta| / lengthscale."""
        raise NotImplementedError

    def grid_columns(self, grid):
        """[*batch, d, gmax]: row i = first Toeplitz column of dimension i (zero padded to gmax)."""
        d = len(grid)
        gmax = max(g.numel() for g in grid)
        ell = self.lengthscale       # [*batch, 1, nd]
        if all(g.numel() == gmax for g in grid):
            # equal grid sizes (every shipped config): one vectorised evaluation over the d dimensions
            G = torch.stack([g.to(ell.dtype) for g in grid])                     # [d, g]
            dist = (G - G[:, :1]).abs()
            li = ell[..., 0, :] if ell.shape[-1] > 1 else ell[..., 0, :].expand(*ell.shape[:-2], d)
            return self.base_1d(dist / li.unsqueeze(-1))                          # [*batch, d, g]
        rows = []
        for i, g in enumerate(grid):
            g = g.to(ell.dtype)
            li = ell[..., 0, i if ell.shape[-1] > 1 else 0].unsqueeze(-1)      # [*batch, 1]
            col = self.base_1d((g - g[0]).abs() / li)                           # [*batch, g_i]
            if g.numel() < gmax:
                col = F.pad(col, (0, gmax - g.numel()))
            rows.append(col)
        return torch.stack(rows, dim=-2)


    def grid_column_dirs(self, grid):
        """d grid_columns / d lengthscale_i per row (no grad), by forward-mode AD through ``base_1d``; ``None`` if the
        grids differ in size.  Only the direction matters (any per-row scaling is absorbed downstream)."""
        import torch.autograd.forward_ad as fwAD
        gmax = max(g.numel() for g in grid)
        if not all(g.numel() == gmax for g in grid):
            return None
        with torch.no_grad():
            ell = self.lengthscale
            d = len(grid)
            G = torch.stack([g.to(ell.dtype) for g in grid])
            dist = (G - G[:, :1]).abs()
            li = (ell[..., 0, :] if ell.shape[-1] > 1 else ell[..., 0, :].expand(*ell.shape[:-2], d)).unsqueeze(-1)
        with torch.enable_grad(), fwAD.dual_level():
            dual = fwAD.make_dual(li.clone(), torch.ones_like(li))
            out = self.base_1d(dist / dual)
            tangent = fwAD.unpack_dual(out).tangent
        return tangent.detach()


class RBFKernel(Kernel):
    def base_1d(self, r):
        return torch.exp(-0.5 * r * r)


class MaternKernel(Kernel):
    def __init__(self, nu=2.5, **kwargs):
        if nu not in {0.5, 1.5, 2.5}:
            raise RuntimeError("nu expected to be 0.5, 1.5, or 2.5")
        super().__init__(**kwargs)
        self.nu = nu

    def base_1d(self, r):
        s = math.sqrt(2 * self.nu) * r
        e = torch.exp(-s)
        if self.nu == 0.5:
            return e
        if self.nu == 1.5:
            return (1 + s) * e
        return (1 + s + 5.0 / 3.0 * r * r) * e


class ScaleKernel(Kernel):
    has_lengthscale = False

    def __init__(self, base_kernel, outputscale_prior=None, outputscale_constraint=None, batch_shape=None, **kwargs):
        bs = base_kernel.batch_shape if batch_shape is None else batch_shape
        super().__init__(batch_shape=bs, **kwargs)
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros(tuple(self.batch_shape)))
        self.raw_outputscale_constraint = outputscale_constraint if outputscale_constraint is not None else Positive()
        if outputscale_prior is not None:
            self._priors["outputscale_prior"] = (outputscale_prior, lambda m: m.outputscale)

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        value = torch.as_tensor(value, dtype=self.raw_outputscale.dtype, device=self.raw_outputscale.device)
        with torch.no_grad():
            self.raw_outputscale.copy_(self.raw_outputscale_constraint.inverse_transform(value).expand_as(self.raw_outputscale))

    def grid_columns(self, grid):
        # last_dim_is_batch=True: every per-dimension factor is scaled (A.3 ii => amplitude outputscale**d)
        cols = self.base_kernel.grid_columns(grid)
        return cols * self.outputscale.reshape(*self.outputscale.shape, 1, 1)

    def grid_column_dirs(self, grid):
        return self.base_kernel.grid_column_dirs(grid)


# ------------------------------------------------------------------ SKI
def create_grid(grid_sizes, grid_bounds, extend=True, device="cpu", dtype=torch.float):
    """``gpytorch.utils.grid.create_grid`` (App. A.2): float32 ``linspace(lo - h, hi + h, g)``, h = (hi-lo)/(g-2)."""
    grid = []
    for i in range(len(grid_bounds)):
        grid_diff = float(grid_bounds[i][1] - grid_bounds[i][0]) / (grid_sizes[i] - 2)
        if extend:
            proj = torch.linspace(float(grid_bounds[i][0] - grid_diff), float(grid_bounds[i][1] + grid_diff),
                                  grid_sizes[i], device=device, dtype=dtype)
        else:
            proj = torch.linspace(float(grid_bounds[i][0]), float(grid_bounds[i][1]), grid_sizes[i], device=device,
                                  dtype=dtype)
        grid.append(proj)
    return grid


class _LazyEvaluatedKernel(LazyTensor):
    """What ``covar_module(X)`` returns: defers to ``evaluate_kernel()`` like GPyTorch's LazyEvaluatedKernelTensor."""

    def __init__(self, kernel, x1, x2=None):
        self.kernel, self.x1, self.x2 = kernel, x1, x2
        self._res = None

    def evaluate_kernel(self):
        if self._res is None:
            self._res = self.kernel._interpolated(self.x1, self.x2)
        return self._res

    def _size(self):
        n2 = self.x1.shape[-2] if self.x2 is None else self.x2.shape[-2]
        return torch.Size((*self.x1.shape[:-2], self.x1.shape[-2], n2))

    def _matmul(self, rhs):
        return self.evaluate_kernel()._matmul(rhs)

    def _transpose_nonbatch(self):
        return _LazyEvaluatedKernel(self.kernel, self.x1 if self.x2 is None else self.x2, self.x1)

    def evaluate(self):
        return self.evaluate_kernel().evaluate()

    def diag(self):
        return self.evaluate_kernel().diag()

    dtype = property(lambda self: self.x1.dtype)
    device = property(lambda self: self.x1.device)


class GridInterpolationKernel(Kernel):
    """SKI kernel  W K_uu W^T  on a regular grid (``gpytorch.kernels.GridInterpolationKernel``)."""
    has_lengthscale = False

    def __init__(self, base_kernel, grid_size, num_dims=None, grid_bounds=None, active_dims=None):
        super().__init__(batch_shape=base_kernel.batch_shape)
        if num_dims is None:
            raise RuntimeError("num_dims must be supplied")
        if grid_bounds is None:
            raise RuntimeError("grid_bounds must be supplied (dynamic grids are not part of the WISKI path)")
        self.base_kernel = base_kernel
        self.num_dims = num_dims
        self.grid_sizes = [grid_size] * num_dims if isinstance(grid_size, int) else [int(g) for g in grid_size]
        self.grid_bounds = grid_bounds
        grid = create_grid(self.grid_sizes, self.grid_bounds)
        for i, g in enumerate(grid):
            self.register_buffer(f"grid_{i}", g)
        self._specs = {}
        self._stencil_memo = None

    @property
    def grid(self):
        return [getattr(self, f"grid_{i}") for i in range(self.num_dims)]

    @property
    def num_inducing(self):
        m = 1
        for g in self.grid_sizes:
            m *= g
        return m

    def grid_spec(self):
        key = self.grid_0.dtype
        if key not in self._specs:
            self._specs[key] = ops.GridSpec(self.grid)
        return self._specs[key]

    def _compute_grid(self, inputs):
        """(idx, val) with the leading batch shape of ``inputs`` kept (A.1 via the interpolation kernel)."""
        batch_shape, n, d = inputs.shape[:-2], inputs.shape[-2], inputs.shape[-1]
        if d != self.num_dims:
            raise RuntimeError(f"expected inputs with {self.num_dims} dimensions, got {d}")
        flat = inputs.reshape(-1, d)
        # one-entry memo: the streaming loop interpolates the same batch twice (evaluate, then condition in update).
        # The memo keeps the tensor alive, so an equal data_ptr + version means unchanged contents.
        memo = self._stencil_memo
        tracked = flat.requires_grad and torch.is_grad_enabled()
        if (memo is not None and not tracked and memo[0].data_ptr() == flat.data_ptr() and memo[0].shape == flat.shape
                and memo[0].dtype == flat.dtype and memo[0].stride() == flat.stride() and memo[1] == flat._version):
            idx, val = memo[2], memo[3]
        else:
            idx, val = ops.interpolate(flat, self.grid_spec(), check_bounds=settings.check_interp_bounds.on())
            self._stencil_memo = None if tracked else (flat.detach(), flat._version, idx, val)
        return idx.view(*batch_shape, n, -1), val.view(*batch_shape, n, -1)

    def _inducing_forward(self, last_dim_is_batch=False, **params):
        """K_uu as a Kronecker product of Toeplitz factors (A.3); one operator per kernel batch element."""
        cols = self.base_kernel.grid_columns(self.grid)
        dirs = None
        if (settings.kron_directional_grad.on() and cols.requires_grad and torch.is_grad_enabled()
                and hasattr(self.base_kernel, "grid_column_dirs")):
            dirs = self.base_kernel.grid_column_dirs(self.grid)
        if cols.dim() == 2:
            return KroneckerToeplitzLazyTensor(cols, self.grid_sizes, dirs)
        from .lazy.lazy_tensor import BatchLazyTensor
        cs = cols.reshape(-1, *cols.shape[-2:])
        ds = [None] * cs.shape[0] if dirs is None else list(dirs.reshape(-1, *dirs.shape[-2:]))
        return BatchLazyTensor([KroneckerToeplitzLazyTensor(c, self.grid_sizes, dd) for c, dd in zip(cs, ds)])

    def _interpolated(self, x1, x2=None):
        li, lv = self._compute_grid(x1)
        if x2 is None or x2 is x1:
            ri, rv = li, lv
        else:
            ri, rv = self._compute_grid(x2)
        # K_uu is built on first use: conditioning and the eval forward only read the stencils
        return InterpolatedLazyTensor(self._inducing_forward, li, lv, ri, rv)

    def forward(self, x1, x2=None, **params):
        if x1.dim() == 1:
            x1 = x1.unsqueeze(-1)
        return _LazyEvaluatedKernel(self, x1, x2)
