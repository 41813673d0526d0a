"""ORACLE (test infrastructure, numpy only).  The integer part of the path — the flat stencil indices of the cubic
SKI interpolation — restated a second time without torch, in scalar-loop form, as an independent check of
``oracle/interp.py`` and of the CUDA kernel (indices must be bit-exact, BASELINE.json north_star).

Follows GPyTorch ``Interpolation.interpolate`` as summarised in SURVEY.md Appendix A.1 (source not under
/root/reference; reference call sites ``online_gp/models/batched_fixed_noise_online_gp.py:143,205,261``):
per dimension ``u = (x - grid[0]) / max(grid[1] - grid[0], eps)`` evaluated in the dtype of x with the grid buffers
read as float32 values, ``lower = floor(u) - 1`` clamped to [0, g - 4] with a one-hot boundary stencil, taps
``lower + {0,1,2,3}`` weighted by the Keys kernel (a = -0.5); flat index in C order with dimension 0 slowest and
stencil column k carrying dimension 0 in its most significant base-4 digit.
"""
import numpy as np


def _keys(s):
    a = abs(s)
    if a < 1:
        return ((1.5 * a - 2.5) * a) * a + 1
    return ((-0.5 * a + 2.5) * a - 4) * a + 2


def interpolate_np(grids, x, eps=1e-10):
    """grids: list of 1-D float32 arrays; x [N, d] (float32 or float64) -> (idx int64 [N, 4^d], val [N, 4^d])."""
    x = np.asarray(x)
    dt = x.dtype.type
    N, d = x.shape
    sizes = [len(g) for g in grids]
    idx = np.zeros((N, 4 ** d), dtype=np.int64)
    val = np.ones((N, 4 ** d), dtype=x.dtype)
    for n in range(N):
        taps, weights = [], []
        for i in range(d):
            g = grids[i]
            lo = dt(g[0])
            delta = dt(max(np.float32(g[1]) - np.float32(g[0]), np.float32(eps)))
            u = dt((dt(x[n, i]) - lo) / delta)
            fl = np.floor(u)
            frac = dt(u - fl)
            lower = int(fl) - 1
            w = [dt(_keys(dt(frac + dt(o)))) for o in (1.0, 0.0, -1.0, -2.0)]
            if lower < 0:
                k = int(np.argmin(np.abs(g[:4].astype(x.dtype) - x[n, i])))
                w = [dt(1.0) if j == k else dt(0.0) for j in range(4)]
                lower = 0
            elif lower > sizes[i] - 4:
                k = int(np.argmin(np.abs(g[-4:].astype(x.dtype) - x[n, i])))
                w = [dt(1.0) if j == k else dt(0.0) for j in range(4)]
                lower = sizes[i] - 4
            taps.append([lower + j for j in range(4)])
            weights.append(w)
        for k in range(4 ** d):
            flat, wv, rem = 0, dt(1.0), k
            digits = []
            for i in range(d):
                digits.append((k // (4 ** (d - 1 - i))) % 4)
            for i in range(d):
                flat = flat * sizes[i] + taps[i][digits[i]]
                wv = dt(wv * weights[i][digits[i]])
            idx[n, k] = flat
            val[n, k] = wv
    return idx, val
