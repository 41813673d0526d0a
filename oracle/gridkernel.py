"""ORACLE (test infrastructure, CPU torch).  Inducing-grid kernel K_uu = kron_i Toeplitz(col_i), restated.

Follows GPyTorch ``GridKernel.forward`` / ``GridInterpolationKernel._inducing_forward`` /
``KroneckerProductLazyTensor._matmul`` / ``ToeplitzLazyTensor._matmul`` and the base kernels
(``RBFKernel``, ``MaternKernel``, ``ScaleKernel`` with ``last_dim_is_batch=True``) as summarised in SURVEY.md
Appendix A.3 / A.4 / A.6 (source not under /root/reference).  Reference call sites:
``online_gp/models/batched_fixed_noise_online_gp.py:107-120`` (default kernel), ``:334-341`` (Kuu / sigma^2),
``:346-348`` (Kuu @ L), ``:366`` (Kuu @ interpolation_cache).

Consequences kept on purpose (A.3): the base kernel is applied per dimension and multiplied (a Matern becomes a
product of 1-D Materns); a ScaleKernel *inside* the GridInterpolationKernel scales every factor, so the effective
amplitude is outputscale**d; ARD lengthscale i applies to factor i; factor i acts on grid axis i (dim 0 slowest).
"""
import math
import torch
import torch.nn.functional as F


class Hypers:
    """Raw (unconstrained) hyper-parameters with GPyTorch's default parametrisation.

    lengthscale = softplus(raw_lengthscale) (Positive), outputscale = softplus(raw_outputscale) (Positive),
    second noise = softplus(raw_noise) + 1e-4 (GreaterThan(1e-4), A.6).  All raw values initialise to 0.
    """

    def __init__(self, d, kind="rbf", has_scale=True, learn_noise=False, dtype=torch.float64):
        self.d, self.kind, self.has_scale, self.learn_noise = d, kind, has_scale, learn_noise
        self.raw_lengthscale = torch.zeros(d, dtype=dtype, requires_grad=True)
        self.raw_outputscale = torch.zeros((), dtype=dtype, requires_grad=True)
        self.raw_noise = torch.zeros((), dtype=dtype, requires_grad=True)

    def params(self):
        p = [self.raw_lengthscale]
        if self.has_scale:
            p.append(self.raw_outputscale)
        if self.learn_noise:
            p.append(self.raw_noise)
        return p

    @property
    def lengthscale(self):
        return F.softplus(self.raw_lengthscale)

    @property
    def outputscale(self):
        return F.softplus(self.raw_outputscale)

    @property
    def noise(self):
        return F.softplus(self.raw_noise) + 1e-4

    def set(self, lengthscale=None, outputscale=None, noise=None):
        inv = lambda v: torch.log(torch.expm1(torch.as_tensor(v, dtype=self.raw_noise.dtype)))
        with torch.no_grad():
            if lengthscale is not None:
                self.raw_lengthscale.copy_(inv(lengthscale).expand(self.d))
            if outputscale is not None:
                self.raw_outputscale.copy_(inv(outputscale))
            if noise is not None:
                self.raw_noise.copy_(inv(torch.as_tensor(noise) - 1e-4))


def base_kernel_1d(kind, dist_over_ell):
    """Stationary base kernel as a function of |delta|/lengthscale (A.3)."""
    r = dist_over_ell
    if kind == "rbf":
        return torch.exp(-0.5 * r * r)
    if kind == "matern0.5":
        return torch.exp(-r)
    if kind == "matern1.5":
        s = math.sqrt(3) * r
        return (1 + s) * torch.exp(-s)
    if kind == "matern2.5":
        s = math.sqrt(5) * r
        return (1 + s + 5.0 / 3.0 * r * r) * torch.exp(-s)
    raise ValueError(kind)


def kuu_columns(grid, hyp):
    """First column of each Toeplitz factor: base_kernel(grid_i[0], grid_i) (* outputscale per factor)."""
    cols = []
    ell = hyp.lengthscale
    for i, gr in enumerate(grid):
        gr = gr.to(ell.dtype)
        col = base_kernel_1d(hyp.kind, (gr - gr[0]).abs() / ell[i])
        if hyp.has_scale:
            col = col * hyp.outputscale
        cols.append(col)
    return cols


def toeplitz_dense(col):
    g = col.shape[0]
    ar = torch.arange(g)
    return col[(ar.unsqueeze(0) - ar.unsqueeze(1)).abs()]


def toeplitz_matmul_fft(col, X, axis):
    """T(col) applied along ``axis`` of X by circulant embedding + FFT — GPyTorch ``ToeplitzLazyTensor._matmul`` ->
    ``toeplitz_matmul`` / ``sym_toeplitz_matmul`` (embedding of length 2g - 1... implemented, as there,
    with the circulant [c_0 .. c_{g-1}, c_{g-1} .. c_1] and real FFTs).  This is what the reference runs per grid axis."""
    g = col.shape[0]
    n = 2 * g - 1 if g > 1 else 1
    circ = torch.cat([col, col[1:].flip(0)]) if g > 1 else col
    fc = torch.fft.rfft(circ, n=n)
    shape = [1] * X.dim()
    fx = torch.fft.rfft(X, n=n, dim=axis)
    shape[axis] = fc.shape[0]
    out = torch.fft.irfft(fx * fc.reshape(shape), n=n, dim=axis)
    return out.narrow(axis, 0, g)


def kron_dense(cols):
    K = torch.ones(1, 1, dtype=cols[0].dtype)
    for c in cols:
        K = torch.kron(K, toeplitz_dense(c))
    return K


def kron_toeplitz_matmul(cols, X):
    """K @ X for X (m x c), factor i acting on grid axis i (dim 0 slowest) — A.4."""
    sizes = [c.shape[0] for c in cols]
    ncol = X.shape[-1]
    Y = X.reshape(*sizes, ncol)
    for i, c in enumerate(cols):
        Y = torch.tensordot(toeplitz_dense(c), Y, dims=([1], [i])).movedim(0, i)
    return Y.reshape(-1, ncol)
