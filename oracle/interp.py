"""ORACLE (test infrastructure, CPU torch).  Cubic-convolution SKI interpolation, restated.

Follows GPyTorch (third-party dependency of the reference, ``requirements.txt:7`` ``gpytorch>=0.3.1``, effective
window 1.3 <= v < 1.9; source NOT under /root/reference) as summarised in SURVEY.md Appendix A.1 / A.2:
``gpytorch.utils.grid.create_grid`` and ``gpytorch.utils.interpolation.Interpolation.interpolate`` /
``left_interp`` / ``InterpolatedLazyTensor._sparse_left_interp_t``.  Reference call sites:
``online_gp/models/batched_fixed_noise_online_gp.py:22-28,143,205-210,261`` and
``online_gp/mlls/streaming_partial_mll.py:20``.
"""
import torch


def create_grid(grid_sizes, grid_bounds, extend=True, dtype=torch.float32):
    """A.2: per dim ``h=(hi-lo)/(g-2)``; ``linspace(lo-h, hi+h, g)`` in **float32** (GPyTorch default)."""
    grid = []
    for g, b in zip(grid_sizes, grid_bounds):
        lo, hi = b[0], b[1]
        diff = float(hi - lo) / (g - 2)
        if extend:
            start, end = lo - diff, hi + diff
        else:
            start, end = lo, hi
        grid.append(torch.linspace(float(start), float(end), int(g), dtype=dtype))
    return grid


def _cubic_kernel(scaled):
    """Keys cubic convolution kernel, a = -0.5 (A.1)."""
    U = scaled.abs()
    lt1 = 1 - U.floor().clamp(0, 1)
    res = (((1.5 * U - 2.5) * U) * U + 1) * lt1
    res = res + (((-0.5 * U + 2.5) * U - 4) * U + 2) * (1 - lt1)
    return res


def interpolate(x_grid, x_target, eps=1e-10):
    """A.1: returns (idx int64 N x 4^d, val N x 4^d); flat index C-order with dim 0 slowest."""
    N, d = x_target.shape
    assert d == len(x_grid)
    sizes = [len(gr) for gr in x_grid]
    for i in range(d):
        gmin, gmax = x_grid[i].min().to(x_target), x_grid[i].max().to(x_target)
        if (x_target[:, i].min() - gmin) < -1e-7 or (x_target[:, i].max() - gmax) > 1e-7:
            raise RuntimeError("Received data that was out of bounds for the specified grid.")
    pts = torch.tensor([-2.0, -1.0, 0.0, 1.0], dtype=x_grid[0].dtype)
    pts_flip = pts.flip(0)
    nc = 4
    vals = torch.ones(N, nc ** d, dtype=x_grid[0].dtype)
    idxs = torch.zeros(N, nc ** d, dtype=torch.long)
    for i in range(d):
        g = sizes[i]
        delta = (x_grid[i][1] - x_grid[i][0]).clamp_min(eps)
        u = (x_target[:, i] - x_grid[i][0]) / delta
        lower = torch.floor(u)
        frac = u - lower
        lower = (lower - pts.max()).detach()
        dvals = _cubic_kernel(frac.unsqueeze(-1) + pts_flip.unsqueeze(-2))
        left = (lower < 0).nonzero().flatten()
        for j in left.tolist():
            first = x_grid[i][:nc].to(x_target)
            k = torch.argmin((first - x_target[j, i]).abs()).item()
            dvals[j, :] = 0
            dvals[j, k] = 1
            lower[j] = 0
        right = (lower > g - nc).nonzero().flatten()
        for j in right.tolist():
            last = x_grid[i][-nc:].to(x_target)
            k = torch.argmin((last - x_target[j, i]).abs()).item()
            dvals[j, :] = 0
            dvals[j, k] = 1
            lower[j] = g - nc
        didx = lower.long().unsqueeze(-1) + torch.arange(nc).unsqueeze(0)
        n_inner, n_outer = nc ** i, nc ** (d - i - 1)
        coeff = 1
        for gs in sizes[i + 1:]:
            coeff *= gs
        didx = didx.unsqueeze(-1).repeat(1, n_inner, n_outer).view(N, -1)
        dvals = dvals.unsqueeze(-1).repeat(1, n_inner, n_outer).view(N, -1)
        idxs = idxs + didx * coeff
        vals = vals * dvals
    return idxs, vals


def left_interp(idx, val, rhs):
    """``W @ rhs``: sum_k val[n,k] * rhs[idx[n,k], :]   (A.1)."""
    return (rhs[idx] * val.unsqueeze(-1)).sum(-2)


def dense_wt(idx, val, m):
    """Dense ``W^T`` (m x N) — what ``_sparse_left_interp_t(...).to_dense()`` yields (duplicates summed)."""
    N = idx.shape[0]
    wt = torch.zeros(m, N, dtype=val.dtype)
    cols = torch.arange(N).unsqueeze(-1).expand_as(idx)
    wt.index_put_((idx.reshape(-1), cols.reshape(-1)), val.reshape(-1), accumulate=True)
    return wt
