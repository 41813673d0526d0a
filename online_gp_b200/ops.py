"""Torch-facing wrappers of the C ABI (device pointers + current stream) and their autograd Functions.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); the arithmetic of every op in this file
runs in libwiski_b200.so.  All tensors must live on a CUDA device — there is no CPU path.
"""
import ctypes
from ctypes import c_double, c_float, c_int, c_int64

import torch

from . import _lib, settings


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("online_gp_b200 ops require CUDA tensors (got a %s tensor); there is no CPU fallback"
                               % t.device.type)


def _sfx(dtype):
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError(f"online_gp_b200: unsupported dtype {dtype} (float32 / float64 only)")


def _real(dtype):
    return c_float if dtype == torch.float32 else c_double


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


#: optional per-op device timing (bench.py): {"names": set(...), "events": {name: [(start, end), ...]}}
PROFILE = None


def _call(name, dtype, *args):
    fn = getattr(_lib.load(), f"{name}_{_sfx(dtype)}")
    _call_fn(name, fn, *args)


def _call_fn(name, fn, *args):
    if PROFILE is not None and name in PROFILE["names"]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(fn(*args), name)
        e1.record()
        PROFILE["events"].setdefault(name, []).append((e0, e1, None))
        return
    _lib.check(fn(*args), name)


class GridSpec:
    """Host-side description of the inducing grid, in the form the interpolation / Kronecker kernels consume.

    Built from the per-dimension grid buffers exactly as GPyTorch's ``Interpolation.interpolate`` reads them
    (SURVEY.md App. A.1): ``lo = grid[0]``, ``delta = (grid[1] - grid[0]).clamp_min(1e-10)`` evaluated in the
    *grid's* dtype (float32 by default, A.2) and only then widened to the working dtype.
    """

    def __init__(self, grid):
        self.d = len(grid)
        if not 1 <= self.d <= 8:
            raise ValueError("grid dimension must be in [1, 8]")
        self.sizes = [int(g.numel()) for g in grid]
        self.m = 1
        for g in self.sizes:
            self.m *= g
        self.s = 4 ** self.d
        self.gmax = max(self.sizes)
        gc = [g.detach().cpu() for g in grid]
        self.lo = [float(g[0]) for g in gc]
        self.delta = [float((g[1] - g[0]).clamp_min(1e-10)) for g in gc]
        self.first4 = [float(v) for g in gc for v in g[:4]]
        self.last4 = [float(v) for g in gc for v in g[-4:]]
        self.gmin = [float(g.min()) for g in gc]
        self.gmax_v = [float(g.max()) for g in gc]
        self.h_sizes = (c_int64 * self.d)(*self.sizes)
        self._host = {}

    def host(self, dtype):
        if dtype not in self._host:
            R = _real(dtype)
            self._host[dtype] = dict(
                lo=(R * self.d)(*self.lo), delta=(R * self.d)(*self.delta), first4=(R * (4 * self.d))(*self.first4),
                last4=(R * (4 * self.d))(*self.last4), gmin=(R * self.d)(*self.gmin), gmax=(R * self.d)(*self.gmax_v))
        return self._host[dtype]


# ---------------------------------------------------------------------------------------------- interpolation
def _interp_fwd(x, spec, check_bounds):
    _require_cuda(x)
    x = x.contiguous()
    q = x.shape[0]
    idx = torch.empty(q, spec.s, dtype=torch.int64, device=x.device)
    val = torch.empty(q, spec.s, dtype=x.dtype, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device) if check_bounds else None
    h = spec.host(x.dtype)
    _call("wiski_interp_fwd", x.dtype, _ptr(x), q, spec.d, spec.h_sizes, h["lo"], h["delta"], h["first4"], h["last4"],
          h["gmin"], h["gmax"], _ptr(idx), _ptr(val), _ptr(flag), _stream())
    if check_bounds and q > 0:
        if _BOUNDS_SINK is not None:
            _BOUNDS_SINK.append((flag, x, spec))          # CUDA-graph capture: the caller reads the flag after replay
        elif settings.defer_interp_bounds_check.on():
            _queue_bounds_check(flag, x, spec)
        elif int(flag.item()) != 0:
            _raise_out_of_bounds(x, spec)
    return idx, val


def _raise_out_of_bounds(x, spec):
    xmin, xmax = x.min().item(), x.max().item()
    raise RuntimeError(
        "Received data that was out of bounds for the specified grid. Grid bounds were (%.3f, %.3f), but min = "
        "%.3f, max = %.3f" % (min(spec.gmin), max(spec.gmax_v), xmin, xmax))


#: set to a list while a CUDA graph is being captured (``OnlineSKIRegression._capture``): flags are collected, not read
_BOUNDS_SINK = None

#: bounds flags whose device->host read was deferred (``settings.defer_interp_bounds_check``): (event, pinned, x, spec)
_PENDING_BOUNDS = []


def _queue_bounds_check(flag, x, spec):
    host = torch.empty(1, dtype=torch.int32, pin_memory=True)
    host.copy_(flag, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _PENDING_BOUNDS.append((ev, host, x, spec))


def flush_bounds_checks():
    """Raise GPyTorch's out-of-bounds RuntimeError for any interpolation whose flag read was deferred.  Waits only
    for the (long finished) flag copies, not for the work queued after them."""
    while _PENDING_BOUNDS:
        ev, host, x, spec = _PENDING_BOUNDS.pop(0)
        ev.synchronize()
        if int(host[0]) != 0:
            _PENDING_BOUNDS.clear()
            _raise_out_of_bounds(x, spec)


class _InterpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, spec, check_bounds):
        idx, val = _interp_fwd(x, spec, check_bounds)
        ctx.save_for_backward(x)
        ctx.spec = spec
        ctx.mark_non_differentiable(idx)
        return idx, val

    @staticmethod
    def backward(ctx, _gidx, gval):
        (x,) = ctx.saved_tensors
        spec = ctx.spec
        x = x.contiguous()
        gval = gval.contiguous()
        gx = torch.empty_like(x)
        h = spec.host(x.dtype)
        _call("wiski_interp_bwd", x.dtype, _ptr(x), x.shape[0], spec.d, spec.h_sizes, h["lo"], h["delta"],
              h["first4"], h["last4"], _ptr(gval), _ptr(gx), _stream())
        return gx, None, None


def interpolate(x, spec, check_bounds=True):
    """x [q,d] -> (idx int64 [q,4^d], val [q,4^d]); differentiable w.r.t. x through val."""
    if x.dim() != 2 or x.shape[1] != spec.d:
        raise ValueError(f"interpolate: expected x of shape [q,{spec.d}], got {tuple(x.shape)}")
    if x.requires_grad and torch.is_grad_enabled():
        return _InterpFn.apply(x, spec, check_bounds)
    return _interp_fwd(x.detach(), spec, check_bounds)


def _gather(idx, val, src):
    _require_cuda(idx, val, src)
    q, s = idx.shape
    m, c = src.shape
    out = torch.empty(q, c, dtype=src.dtype, device=src.device)
    _call("wiski_gather", src.dtype, _ptr(idx), _ptr(val), q, s, _ptr(src), m, c, _ptr(out), _stream())
    return out


def _scatter_add(idx, val, src, dst):
    _require_cuda(idx, val, src, dst)
    q, s = idx.shape
    m, c = dst.shape
    _call("wiski_scatter_add", dst.dtype, _ptr(idx), _ptr(val), q, s, _ptr(src), m, c, _ptr(dst), _stream())
    return dst


class _LeftInterpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, val, src):
        ctx.save_for_backward(idx, val, src)
        return _gather(idx, val.contiguous(), src.contiguous())

    @staticmethod
    def backward(ctx, gout):
        idx, val, src = ctx.saved_tensors
        gout = gout.contiguous()
        gval = gsrc = None
        if ctx.needs_input_grad[1]:
            gval = (src[idx] * gout.unsqueeze(1)).sum(-1)
        if ctx.needs_input_grad[2]:
            gsrc = torch.zeros_like(src)
            _scatter_add(idx, val.contiguous(), gout, gsrc)
        return None, gval, gsrc


def left_interp(idx, val, src):
    """W @ src: out[n,:] = sum_k val[n,k] src[idx[n,k],:]  (src [m,c] -> [q,c])."""
    if src.dim() != 2:
        raise ValueError("left_interp: src must be [m,c]")
    if torch.is_grad_enabled() and (val.requires_grad or src.requires_grad):
        return _LeftInterpFn.apply(idx, val, src)
    return _gather(idx, val.detach().contiguous(), src.detach().contiguous())


def left_t_interp(idx, val, src, m):
    """W^T @ src: [m,c] from src [q,c] (no autograd; used for cache accumulation under no_grad)."""
    dst = torch.zeros(m, src.shape[-1], dtype=src.dtype, device=src.device)
    return _scatter_add(idx, val.detach().contiguous(), src.detach().contiguous(), dst)


def scatter_add_(dst, idx, val, src):
    return _scatter_add(idx, val.detach().contiguous(), src.detach().contiguous(), dst)


# ---------------------------------------------------------------------------------------------- Kronecker-Toeplitz
def _tc_axis_work(X, g, outer, inner, contract):
    """Scratch for the tensor-core form of one Kronecker axis (fp32, g >= 64), or None if the shape is not eligible."""
    if X.dtype != torch.float32 or not X.is_cuda or g < 64 or not settings.kron_tensor_core_axes.on():
        return None
    n = int(_lib.load().wiski_kron_axis_tc_work_elems(int(g), int(outer), int(inner), int(contract)))
    return torch.empty(n, dtype=torch.float32, device=X.device) if n > 0 else None


def _kron_mm(cols, sizes, X):
    _require_cuda(cols, X)
    d, gmax = cols.shape
    m, c = X.shape
    X = X.contiguous()
    if X.dtype == torch.float32 and X.is_cuda and max(sizes) >= 64 and settings.kron_tensor_core_axes.on():
        # large axes: per-axis passes, each on the tensor cores when its shape is eligible (kron_axis_apply)
        for (g, outer, inner), col in zip(_axis_geometry(sizes, c), cols):
            X = kron_axis_apply(X, col, g, outer, inner)
        return X
    Y = torch.empty_like(X)
    work = torch.empty_like(X) if d > 1 else None
    h_g = (c_int64 * d)(*sizes)
    _call("wiski_kron_toeplitz_mm", X.dtype, _ptr(cols), d, h_g, gmax, _ptr(X), c, _ptr(Y), _ptr(work), _stream())
    return Y


def _kron_bwd_cols(cols, sizes, Z, X):
    d, gmax = cols.shape
    m, c = X.shape
    esz = X.element_size()
    n = _lib.load().wiski_kron_toeplitz_bwd_work_elems(d, m, c, gmax, esz)
    work = torch.empty(n, dtype=X.dtype, device=X.device)
    g = torch.empty_like(cols)
    h_g = (c_int64 * d)(*sizes)
    _call("wiski_kron_toeplitz_bwd_cols", X.dtype, _ptr(cols), d, h_g, gmax, _ptr(Z.contiguous()), _ptr(X.contiguous()),
          c, _ptr(g), _ptr(work), _stream())
    return g


def _axis_geometry(sizes, c):
    """[(g_i, outer_i, inner_i)] for a [prod(sizes), c] panel viewed as [outer, g_i, inner]."""
    m = 1
    for g in sizes:
        m *= g
    out, outer = [], 1
    for g in sizes:
        out.append((g, outer, (m // (outer * g)) * c))
        outer *= g
    return out


def _fused_supported(sizes, X):
    """Two-axes-per-pass fast path (kron_fused.cu): fp32, every grid axis 32 points, d even, c % 16 == 0."""
    if X.dtype != torch.float32 or not X.is_cuda:
        return False
    h_g = (c_int64 * len(sizes))(*sizes)
    return bool(_lib.load().wiski_kron_fused_supported(len(sizes), h_g, X.shape[1]))


def _pair_axes(pair):
    """pair index p -> grid axes (2p, 2p + 1); an explicit (axis_u, axis_v) tuple passes through."""
    return (2 * pair, 2 * pair + 1) if isinstance(pair, int) else (int(pair[0]), int(pair[1]))


def kron_pairs(sizes, directional=True):
    """How the grid axes are paired for the fused two-axes passes (applied last to first).  4-D grids on the tensor-core
    path use (0,3) + (1,2) — see settings.kron_outer_inner_pairing; the SIMT full-gradient pass only knows adjacent pairs."""
    d = len(sizes)
    if (d == 4 and directional and settings.kron_outer_inner_pairing.on() and settings.kron_directional_grad.on()
            and _lib.load().wiski_kron_tc_enable(-1)):
        return [(0, 3), (1, 2)]
    return [(2 * p, 2 * p + 1) for p in range(d // 2)]


def _lay_array(c, *chunked):
    """(ld, cw, cstride) per operand for the *_lay entry points; `chunked` item = None (plain [m, c]) or (W, m):
    column-chunked [W, m, c / W] (block j = columns [j c/W, (j+1) c/W), row-major with pitch c / W)."""
    vals = []
    for ch in chunked:
        if ch is None:
            vals += [c, c, 0]
        else:
            W, m = ch
            vals += [c // W, c // W, m * (c // W)]
    return (c_int64 * len(vals))(*[int(v) for v in vals])


def _fused_pair_apply(cols, sizes, pair, X, chunk_out=1, out=None):
    """Y = (T_u x T_v) X for X [m, c]; pair = index p (axes 2p, 2p+1) or an (axis_u, axis_v) tuple.  chunk_out = W > 1: Y is
    written column-chunked, [W, m, c / W] (the send buffer of the row -> column all-to-all of the sharded path: no
    transposing copy)."""
    d, gmax = cols.shape
    m, c = X.shape
    h_g = (c_int64 * d)(*sizes)
    au, av = _pair_axes(pair)
    fn = _lib.load().wiski_kron_pair_apply_axes_f32
    if chunk_out > 1:
        if c % (16 * chunk_out) != 0:
            raise ValueError("column-chunked output needs c % (16 * chunks) == 0")
        Y = torch.empty(chunk_out, m, c // chunk_out, dtype=X.dtype, device=X.device) if out is None else out
        if Y.shape != (chunk_out, m, c // chunk_out) or not Y.is_contiguous() or Y.dtype != X.dtype:
            raise ValueError("_fused_pair_apply: bad `out`")
        _call_fn("wiski_kron_fused_pair_apply", fn, _ptr(cols), d, h_g, gmax, au, av, _ptr(X), _ptr(Y), c,
                 _lay_array(c, None, (chunk_out, m)), _stream())
        return Y
    Y = torch.empty_like(X) if out is None else out
    if Y.shape != X.shape or not Y.is_contiguous() or Y.dtype != X.dtype or Y.data_ptr() == X.data_ptr():
        raise ValueError("_fused_pair_apply: bad `out`")
    _call_fn("wiski_kron_fused_pair_apply", fn, _ptr(cols), d, h_g, gmax, au, av, _ptr(X), _ptr(Y), c, None, _stream())
    return Y


def _fused_pair_grad(cols, sizes, pair, Z, P, acc, store, chunk_z=1, zout=None):
    """One backward pair pass (wiski_kron_fused_pair_grad).  chunk_z = W > 1: Z is given column-chunked [W, m, c / W]
    (as received from the column -> row all-to-all); P and the returned Zout are plain [m, c]."""
    d, gmax = cols.shape
    m, c = P.shape
    Zout = (torch.empty_like(P) if zout is None else zout) if store else None
    if Zout is not None and (Zout.shape != P.shape or not Zout.is_contiguous() or Zout.data_ptr() in (Z.data_ptr(), P.data_ptr())):
        raise ValueError("_fused_pair_grad: bad `zout`")
    h_g = (c_int64 * d)(*sizes)
    if chunk_z > 1:
        _call_fn("wiski_kron_fused_pair_grad", _lib.load().wiski_kron_fused_pair_grad_lay_f32, _ptr(cols), d, h_g, gmax,
                 pair, _ptr(Z), _ptr(P), _ptr(Zout), c, _ptr(acc[2 * pair]), _ptr(acc[2 * pair + 1]),
                 _lay_array(c, (chunk_z, m), None, None), _stream())
        return Zout
    _call_fn("wiski_kron_fused_pair_grad", _lib.load().wiski_kron_fused_pair_grad_f32, _ptr(cols), d, h_g, gmax, pair,
             _ptr(Z), _ptr(P), _ptr(Zout), c, _ptr(acc[2 * pair]), _ptr(acc[2 * pair + 1]), _stream())
    return Zout


def _fused_pair_grad_dir(cols, dirs, sizes, pair, Z, P, out3, store, chunk_z=1, zout=None):
    """One directional backward pair pass: out3 += [<g_u, dirs_u>, <g_v, dirs_v>, <Z', K' P'>] for the axes of `pair` (index p
    or (axis_u, axis_v)).  chunk_z = W > 1: Z is given column-chunked [W, m, c / W] (received from the column -> row
    all-to-all); P and the returned Zout are plain [m, c]."""
    d, gmax = cols.shape
    m, c = P.shape
    Zout = (torch.empty_like(P) if zout is None else zout) if store else None
    if Zout is not None and (Zout.shape != P.shape or not Zout.is_contiguous() or Zout.data_ptr() in (Z.data_ptr(), P.data_ptr())):
        raise ValueError("_fused_pair_grad_dir: bad `zout`")
    h_g = (c_int64 * d)(*sizes)
    au, av = _pair_axes(pair)
    lay = _lay_array(c, (chunk_z, m), None, None) if chunk_z > 1 else None
    _call_fn("wiski_kron_fused_pair_grad", _lib.load().wiski_kron_pair_grad_dir_axes_f32, _ptr(cols), _ptr(dirs), d, h_g, gmax,
             au, av, _ptr(Z), _ptr(P), _ptr(Zout), c, _ptr(out3), lay, _stream())
    return Zout


# ---- pushing variants (row-sharded multi-GPU path): the result leaves through the kernel's own stores into the peers'
# NVLink-mapped buffers; `dst` = ctypes table of device pointers (one per rank, parallel._PushBuffers.dst)
def _fused_pair_apply_push(cols, sizes, pair, X, dst, n_dst, mode):
    """(T_u x T_v) X pushed to the peers.  mode 1: X is a row slab [m_loc, c], column block j -> rank j (a [m_loc, c / n_dst]
    panel at dst[j]); mode 2: X holds all rows of a column block, axis_u = 0, rows of axis-0 range j -> rank j; mode 3: mode 2
    with 32-column store boxes (c % 32 == 0)."""
    d, gmax = cols.shape
    m, c = X.shape
    h_g = (c_int64 * d)(*sizes)
    au, av = _pair_axes(pair)
    _call_fn("wiski_kron_fused_pair_apply", _lib.load().wiski_kron_pair_apply_push_f32, _ptr(cols), d, h_g, gmax, au, av,
             _ptr(X), c, None, dst, int(n_dst), int(mode), _stream())


def _fused_pair_grad_dir_push(cols, dirs, sizes, pair, Z, P, out3, dst, n_dst):
    """Directional backward pair pass on plain [m, c] operands (axis_u = 0) whose Zout = T_v T_u Z is pushed to the peers:
    rows of axis-0 range j -> rank j (a [m / n_dst, c] panel at dst[j])."""
    d, gmax = cols.shape
    m, c = P.shape
    h_g = (c_int64 * d)(*sizes)
    au, av = _pair_axes(pair)
    _call_fn("wiski_kron_fused_pair_grad", _lib.load().wiski_kron_pair_grad_dir_push_f32, _ptr(cols), _ptr(dirs), d, h_g, gmax,
             au, av, _ptr(Z), _ptr(P), c, _ptr(out3), None, dst, int(n_dst), _stream())


def rmul_push(P, M, dst, n_dst, terms=3):
    """P M with column block j (r2 / n_dst columns) written at dst[j] (a [m, r2 / n_dst] panel on rank j); no autograd."""
    _require_cuda(P, M)
    P, M = P.contiguous(), M.contiguous()
    m, r = P.shape
    r2 = M.shape[1]
    _call_fn("wiski_panel_rmul", _lib.load().wiski_panel_rmul_push_f32, _ptr(P), m, r, _ptr(M), r2, int(terms), dst, int(n_dst),
             _ptr(_rmul_work(r, r2, terms, P.device)), _stream())


def _by_axis(out, pairs, d):
    """out [npairs, 3] (per pair: axis_u, axis_v, scale) -> the d directional sums ordered by grid axis."""
    order = [0] * d
    for p, (au, av) in enumerate(pairs):
        order[au], order[av] = 2 * p, 2 * p + 1
    flat = out[:, :2].reshape(-1)
    return torch.stack([flat[i] for i in order])        # (no index tensor: nothing may be copied from the host under capture)


def _surrogate_col_grad(cols, dirs, s_dir, s_scale):
    """The vector g_i in span{dirs_i, cols_i} with <g_i, dirs_i> = s_dir_i and <g_i, cols_i> = s_scale: equivalent to
    the true column gradient for every parameter that moves col_i only along those two directions."""
    c64, d64 = cols.double(), dirs.double()
    dd, dc, cc = (d64 * d64).sum(-1), (d64 * c64).sum(-1), (c64 * c64).sum(-1)
    det = dd * cc - dc * dc
    alpha = (s_dir * cc - s_scale * dc) / det
    beta = (dd * s_scale - dc * s_dir) / det
    return (alpha.unsqueeze(-1) * d64 + beta.unsqueeze(-1) * c64).to(cols.dtype)


class _KronFn(torch.autograd.Function):
    """K X with the column gradient.  When cols needs grad the forward applies the axes in the order d-1, ..., 0 and
    keeps the suffix products S_i = T_{i+1} .. T_{d-1} X, so the backward only runs the prefix chain on the incoming
    gradient plus one contraction per axis (d-1 axis passes instead of 2(d-1))."""

    #: optional destination of the next forward's result (set by ``kron_toeplitz_matmul(out=...)``; kept out of the
    #: autograd inputs so that the result is an ordinary fresh output as far as autograd is concerned)
    _next_out = None

    @staticmethod
    def forward(ctx, cols, X, sizes, dirs=None):
        out, _KronFn._next_out = _KronFn._next_out, None
        cols = cols.contiguous()
        ctx.sizes = sizes
        ctx.dirs = None
        if not ctx.needs_input_grad[0]:
            ctx.save_for_backward(cols, X)
            ctx.suffix = None
            Y = _kron_mm(cols, sizes, X)
            return Y if out is None else out.copy_(Y)
        if _fused_supported(sizes, X):
            # pairs applied last to first; M[p] = (pairs > p) applied to X is what the backward pair pass p needs
            pairs = kron_pairs(sizes, directional=dirs is not None)
            npairs = len(pairs)
            M = [None] * npairs
            M[-1] = X.contiguous()
            for p in range(npairs - 1, 0, -1):
                M[p - 1] = _fused_pair_apply(cols, sizes, pairs[p], M[p])
            Y = _fused_pair_apply(cols, sizes, pairs[0], M[0], out=out)       # `out`: e.g. the send buffer of an exchange
            ctx.save_for_backward(cols, *M)
            ctx.suffix = "fused"
            ctx.pairs = pairs
            ctx.dirs = None if dirs is None else dirs.detach().to(cols.dtype).contiguous()
            return Y
        geo = _axis_geometry(sizes, X.shape[1])
        S = [None] * len(sizes)
        S[-1] = X.contiguous()
        for i in range(len(sizes) - 1, 0, -1):
            g, outer, inner = geo[i]
            S[i - 1] = kron_axis_apply(S[i], cols[i], g, outer, inner)
        g, outer, inner = geo[0]
        Y = kron_axis_apply(S[0], cols[0], g, outer, inner)
        ctx.save_for_backward(cols, *S)
        ctx.suffix = True
        return Y if out is None else out.copy_(Y)

    @staticmethod
    def backward(ctx, gY):
        gY = gY.contiguous()
        if ctx.suffix is None:
            cols, X = ctx.saved_tensors
            return None, (_kron_mm(cols, ctx.sizes, gY) if ctx.needs_input_grad[1] else None), None, None
        cols, *S = ctx.saved_tensors
        sizes = ctx.sizes
        d, gmax = cols.shape
        acc = torch.zeros(d, gmax, dtype=torch.float64, device=gY.device)
        if ctx.suffix == "fused":
            Zc = gY
            pairs = ctx.pairs
            npairs = len(pairs)
            if ctx.dirs is not None:
                # directional form: direction-matrix applies + dot products instead of the 32-entry column gradients
                out = torch.zeros(npairs, 3, dtype=torch.float64, device=gY.device)
                for p in range(npairs):
                    Zc = _fused_pair_grad_dir(cols, ctx.dirs, sizes, pairs[p], Zc, S[p], out[p],
                                              store=(p < npairs - 1 or ctx.needs_input_grad[1]))
                gcols = _surrogate_col_grad(cols[:, :sizes[0]], ctx.dirs[:, :sizes[0]], _by_axis(out, pairs, d), out[-1, 2])
                if gcols.shape[1] < gmax:
                    gcols = torch.nn.functional.pad(gcols, (0, gmax - gcols.shape[1]))
                return gcols, (Zc if ctx.needs_input_grad[1] else None), None, None
            for p in range(npairs):
                Zc = _fused_pair_grad(cols, sizes, p, Zc, S[p], acc, store=(p < npairs - 1 or ctx.needs_input_grad[1]))
            return acc.to(cols.dtype), (Zc if ctx.needs_input_grad[1] else None), None, None
        geo = _axis_geometry(sizes, gY.shape[1])
        Pz = gY
        for i in range(d):
            g, outer, inner = geo[i]
            kron_axis_contract(Pz, S[i], g, outer, inner, acc[i])
            if i < d - 1 or ctx.needs_input_grad[1]:
                Pz = kron_axis_apply(Pz, cols[i], g, outer, inner)
        gX = Pz if ctx.needs_input_grad[1] else None      # K symmetric: after all d axes Pz = K gY
        return acc.to(cols.dtype), gX, None, None


def kron_toeplitz_matmul(cols, sizes, X, dirs=None, out=None):
    """(T(cols[0]) x ... x T(cols[d-1])) @ X for X [m,c]; cols [d,gmax] (row i valid in its first sizes[i] entries).
    Differentiable w.r.t. cols and X.  ``dirs`` [d,gmax] (optional): d cols[i] / d lengthscale_i; when given (and the
    fused fp32 path applies) the backward returns a *surrogate* column gradient that is exact for parameters moving
    cols[i] along dirs[i] or along cols[i] itself (lengthscales and scalar scales) — see ``_surrogate_col_grad``."""
    if X.dim() != 2:
        raise ValueError("kron_toeplitz_matmul: X must be [m,c]")
    if cols.dtype != X.dtype:
        raise TypeError("kron_toeplitz_matmul: cols and X must share a dtype")
    if torch.is_grad_enabled() and (cols.requires_grad or X.requires_grad):
        _KronFn._next_out = out
        try:
            return _KronFn.apply(cols, X, tuple(sizes), dirs)
        finally:
            _KronFn._next_out = None
    Y = _kron_mm(cols.detach().contiguous(), tuple(sizes), X.detach())
    return Y if out is None else out.copy_(Y)


def kron_axis_apply(X, col, g, outer, inner):
    """One Kronecker axis: view X as [outer, g, inner]; Y[o,a,w] = sum_b col[|a-b|] X[o,b,w]. (no autograd)"""
    _require_cuda(X, col)
    X = X.contiguous()
    Y = torch.empty_like(X)
    work = _tc_axis_work(X, g, outer, inner, 0)
    if work is not None:
        _call_fn("wiski_kron_axis_apply", _lib.load().wiski_kron_axis_apply_tc_f32, _ptr(X), _ptr(Y),
                 _ptr(col.contiguous()), int(g), int(outer), int(inner), _ptr(work), _stream())
        return Y
    _call("wiski_kron_axis_apply", X.dtype, _ptr(X), _ptr(Y), _ptr(col.contiguous()), int(g), int(outer), int(inner),
          _stream())
    return Y


def kron_axis_contract(Z, P, g, outer, inner, acc64):
    """acc64[k] += sum over lines of sum_{|a-b|=k} Z[o,a,w] P[o,b,w]  (acc64: float64 [g], accumulated in place)."""
    _require_cuda(Z, P, acc64)
    work = _tc_axis_work(Z, g, outer, inner, 1)
    if work is not None:
        _call_fn("wiski_kron_axis_contract", _lib.load().wiski_kron_axis_contract_tc_f32, _ptr(Z.contiguous()),
                 _ptr(P.contiguous()), int(g), int(outer), int(inner), _ptr(acc64), _ptr(work), _stream())
        return acc64
    _call("wiski_kron_axis_contract", Z.dtype, _ptr(Z.contiguous()), _ptr(P.contiguous()), int(g), int(outer), int(inner),
          _ptr(acc64), _stream())
    return acc64


# ---------------------------------------------------------------------------------------------- panels
def _rmul_work(r, r2, terms, device):
    if terms != 2:
        return None
    return torch.empty(int(_lib.load().wiski_panel_rmul_ex_work_elems(r, r2)), dtype=torch.float32, device=device)


def _rmul(P, M, terms=3):
    """terms: tcgen05 passes of the fp32 tensor-core path (3 = 3xTF32 split, fp32-grade; 1 = single tf32 pass,
    gradient quantities only — settings.backward_gemm_tf32_passes)."""
    _require_cuda(P, M)
    P, M = P.contiguous(), M.contiguous()
    m, r = P.shape
    r2 = M.shape[1]
    out = torch.empty(m, r2, dtype=P.dtype, device=P.device)
    if terms in (1, 2) and P.dtype == torch.float32:
        _call_fn("wiski_panel_rmul", _lib.load().wiski_panel_rmul_ex_f32, _ptr(P), m, r, _ptr(M), r2, 1, int(terms), _ptr(out),
                 _ptr(_rmul_work(r, r2, terms, P.device)), _stream())
        return out
    _call("wiski_panel_rmul", P.dtype, _ptr(P), m, r, _ptr(M), r2, _ptr(out), _stream())
    return out


def _gram(A, B, symmetric=False):
    """symmetric: the caller knows A^T B is symmetric (Q - I = L^T (K L)): the fp32 tensor-core path then skips the
    result tiles below the diagonal and mirrors them."""
    _require_cuda(A, B)
    A, B = A.contiguous(), B.contiguous()
    m, r = A.shape
    r2 = B.shape[1]
    n = _lib.load().wiski_gram_work_elems(m, r, r2)
    work = torch.empty(max(int(n), 1), dtype=A.dtype, device=A.device)
    G = torch.empty(r, r2, dtype=A.dtype, device=A.device)
    if symmetric and r == r2 and A.dtype == torch.float32:
        _call_fn("wiski_gram", _lib.load().wiski_gram_sym_f32, _ptr(A), _ptr(B), m, r, _ptr(G), _ptr(work), _stream())
        return G
    _call("wiski_gram", A.dtype, _ptr(A), _ptr(B), m, r, r2, _ptr(G), _ptr(work), _stream())
    return G


def gram_blocks(A, Bb, symmetric=False):
    """A^T [B_0 | B_1 | ...] for column blocks Bb [nb, m, cwb] -> [r, nb cwb] (no autograd)."""
    _require_cuda(A, Bb)
    nb, m, cwb = Bb.shape
    r, r2 = A.shape[1], nb * cwb
    if nb == 1:
        return _gram(A, Bb[0], symmetric)
    if A.dtype == torch.float32 and cwb % 32 == 0:
        A, Bb = A.contiguous(), Bb.contiguous()
        lib = _lib.load()
        work = torch.empty(max(int(lib.wiski_gram_work_elems(m, r, r2)), 1), dtype=A.dtype, device=A.device)
        G = torch.empty(r, r2, dtype=A.dtype, device=A.device)
        if symmetric and r == r2:
            rc = lib.wiski_gram_chunked_sym_f32(_ptr(A), _ptr(Bb), m, r, nb, _ptr(G), _ptr(work), _stream())
        else:
            rc = lib.wiski_gram_chunked_f32(_ptr(A), _ptr(Bb), m, r, r2, nb, _ptr(G), _ptr(work), _stream())
        if rc == 0:
            return G
        if rc != 3:
            _lib.check(rc, "wiski_gram_chunked")
    return torch.cat([_gram(A, Bb[j]) for j in range(nb)], dim=1)


def rmul_blocks(P, M, nb, out=None, terms=3):
    """[Out_0 | Out_1 | ...] = P M returned as column blocks [nb, m, r2 / nb] (no autograd)."""
    _require_cuda(P, M)
    m, r = P.shape
    r2 = M.shape[1]
    cwb = r2 // nb
    if nb == 1:
        return _rmul(P, M, terms).unsqueeze(0)
    if P.dtype == torch.float32 and cwb % 32 == 0:
        P, M = P.contiguous(), M.contiguous()
        Out = torch.empty(nb, m, cwb, dtype=P.dtype, device=P.device) if out is None else out
        rc = _lib.load().wiski_panel_rmul_ex_f32(_ptr(P), m, r, _ptr(M), r2, nb, int(terms), _ptr(Out),
                                                 _ptr(_rmul_work(r, r2, terms, P.device)), _stream())
        if rc == 0:
            return Out
        if rc != 3:
            _lib.check(rc, "wiski_panel_rmul_chunked")
    res = torch.stack([_rmul(P, M[:, j * cwb:(j + 1) * cwb].contiguous(), terms) for j in range(nb)])
    return res if out is None else out.copy_(res)


def _bwd_terms():
    return int(settings.backward_gemm_tf32_passes.value())


class _RmulFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, P, M):
        ctx.save_for_backward(P, M)
        return _rmul(P, M)

    @staticmethod
    def backward(ctx, gO):
        P, M = ctx.saved_tensors
        gP = gM = None
        if ctx.needs_input_grad[0]:
            gP = _rmul(gO, M.t().contiguous(), _bwd_terms())
        if ctx.needs_input_grad[1]:
            gM = _gram(P, gO)
        return gP, gM


class _GramFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, symmetric=False):
        ctx.save_for_backward(A, B)
        ctx.bg = is_background()         # forward ran as a background pass on a side stream: so does the backward
        return _gram(A, B, symmetric)

    @staticmethod
    def backward(ctx, gG):
        A, B = ctx.saved_tensors
        gA = gB = None
        with background(ctx.bg):
            if ctx.needs_input_grad[0]:
                gA = _rmul(B, gG.t().contiguous(), _bwd_terms())
            if ctx.needs_input_grad[1]:
                gB = _rmul(A, gG.contiguous(), _bwd_terms())
        return gA, gB, None


def panel_rmul(P, M):
    """P [m,r] @ M [r,r2] -> [m,r2]."""
    if torch.is_grad_enabled() and (P.requires_grad or M.requires_grad):
        return _RmulFn.apply(P, M)
    return _rmul(P.detach(), M.detach())


def gram(A, B, symmetric=False):
    """A^T @ B for panels A [m,r], B [m,r2] -> [r,r2].  ``symmetric``: the caller knows the result is symmetric."""
    if torch.is_grad_enabled() and (A.requires_grad or B.requires_grad):
        return _GramFn.apply(A, B, symmetric)
    return _gram(A.detach(), B.detach(), symmetric)


def panel_lowrank_update_(P, U, Vt):
    """In place: P <- P + (P @ U) @ Vt with U [r,q], Vt [q,r], q <= 32."""
    _require_cuda(P, U, Vt)
    if not P.is_contiguous():
        raise ValueError("panel_lowrank_update_: P must be contiguous")
    m, r = P.shape
    q = U.shape[1]
    _call("wiski_panel_lowrank_update", P.dtype, _ptr(P), m, r, _ptr(U.contiguous()), _ptr(Vt.contiguous()), q, _stream())
    return P


def panel_lowrank_update2_(P0, P1, U, Vt0, Vt1, return_t=False):
    """In place, one launch: P0 <- P0 + (P0 @ U) @ Vt0 and P1 <- P1 + (P1 @ U) @ Vt1 (the root / inverse-root pair of the
    rank-q update; U [r,q] shared, Vt [q,r]).  q > 32 is applied in column blocks of 32 — exact, because
    I + U Vt is only ever used here with Vt = C U^T blocks that the callers already chunk; see ``_apply``."""
    _require_cuda(P0, P1, U, Vt0, Vt1)
    if not (P0.is_contiguous() and P1.is_contiguous()) or P0.shape != P1.shape:
        raise ValueError("panel_lowrank_update2_: panels must be contiguous and of equal shape")
    m, r = P0.shape
    q = U.shape[1]
    if return_t:            # also hand back P0 @ U of the rows before the update (by-product of the same pass)
        T = torch.empty(m, q, dtype=P0.dtype, device=P0.device)
        _call("wiski_panel_lowrank_update2_t", P0.dtype, _ptr(P0), _ptr(P1), m, r, _ptr(U.contiguous()), _ptr(Vt0.contiguous()),
              _ptr(Vt1.contiguous()), q, _ptr(T), _stream())
        return P0, P1, T
    _call("wiski_panel_lowrank_update2", P0.dtype, _ptr(P0), _ptr(P1), m, r, _ptr(U.contiguous()), _ptr(Vt0.contiguous()),
          _ptr(Vt1.contiguous()), q, _stream())
    return P0, P1


class background:
    """Context manager: library launches inside run in "background" form (small resident footprint, largest shared-memory
    carve-out; ``wiski_set_background``) — for HBM-bound passes put on a side stream under a tensor-bound kernel."""

    def __init__(self, on=True):
        self.on = bool(on)

    def __enter__(self):
        self.prev = _lib.load().wiski_set_background(1 if self.on else 0)
        return self

    def __exit__(self, *exc):
        _lib.load().wiski_set_background(self.prev)
        return False


def is_background():
    return bool(_lib.load().wiski_set_background(-1))


_SIDE_STREAMS = {}


def side_stream(device):
    """The one side stream per device used by the overlap schedules (settings.overlap_root_update)."""
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def overlap_capable(t):
    """Side-stream overlap applies to CUDA tensors only (a hook the CPU tests replace to exercise the schedule's logic)."""
    return bool(t.is_cuda)


class side_section:
    """``with side_section(device):`` — fork the side stream from the current stream and run the body on it as background
    launches; ``join_side(device)`` makes the current stream wait for it.  Stream events only, so both are capturable."""

    def __init__(self, device):
        self.device = device

    def __enter__(self):
        side = side_stream(self.device)
        side.wait_stream(torch.cuda.current_stream())
        self._stream_ctx = torch.cuda.stream(side)
        self._stream_ctx.__enter__()
        self._bg = background()
        self._bg.__enter__()
        return self

    def __exit__(self, *exc):
        self._bg.__exit__(*exc)
        self._stream_ctx.__exit__(*exc)
        return False


def join_side(device):
    torch.cuda.current_stream().wait_stream(side_stream(device))


def panel_lowrank_update1_(P, U, Vt, return_t=False):
    """In place: P <- P + (P @ U) @ Vt through the several-rows-per-warp kernel of ``panel_lowrank_update2_`` (one panel);
    return_t: also P @ U of the rows before the update."""
    _require_cuda(P, U, Vt)
    if not P.is_contiguous():
        raise ValueError("panel_lowrank_update1_: panel must be contiguous")
    m, r = P.shape
    if return_t:
        T = torch.empty(m, U.shape[1], dtype=P.dtype, device=P.device)
        _call("wiski_panel_lowrank_update2_t", P.dtype, _ptr(P), _ptr(None), m, r, _ptr(U.contiguous()), _ptr(Vt.contiguous()),
              _ptr(None), U.shape[1], _ptr(T), _stream())
        return P, T
    _call("wiski_panel_lowrank_update2", P.dtype, _ptr(P), _ptr(None), m, r, _ptr(U.contiguous()), _ptr(Vt.contiguous()),
          _ptr(None), U.shape[1], _stream())
    return P


def panel_outer_add_(P, T, W):
    """In place, one streaming pass: P [m,c] += T [m,q] @ W [q,c] (q <= 32)."""
    _require_cuda(P, T, W)
    if not P.is_contiguous() or T.shape[0] != P.shape[0] or W.shape != (T.shape[1], P.shape[1]):
        raise ValueError("panel_outer_add_: P [m,c] contiguous, T [m,q], W [q,c]")
    m, c = P.shape
    _call("wiski_panel_outer_add", P.dtype, _ptr(P), m, c, _ptr(T.contiguous()), T.shape[1], _ptr(W.contiguous()), _stream())
    return P


def q_matvec(L, KL, v):
    """w = v + L^T (KL v) for v [r,c] (c <= 4 per launch; wider right-hand sides are split)."""
    _require_cuda(L, KL, v)
    m, r = L.shape
    v = v.contiguous()
    outs = []
    lib = _lib.load()
    for c0 in range(0, v.shape[1], 4):
        vc = v[:, c0:c0 + 4].contiguous()
        c = vc.shape[1]
        work = torch.empty(int(lib.wiski_qmv_work_elems(m, r, c)), dtype=L.dtype, device=L.device)
        w = torch.empty_like(vc)
        _call("wiski_q_matvec", L.dtype, _ptr(L), _ptr(KL), m, r, _ptr(vc), c, _ptr(w), _ptr(work), _stream())
        outs.append(w)
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


def cg_solve(L, KL, rhs, tol=1e-2, max_iter=1000, check_every=4):
    """Solve (I + L^T KL) x = rhs by CG with the fused panel MVM. Returns (x, iters, residual)."""
    _require_cuda(L, KL, rhs)
    m, r = L.shape
    rhs = rhs.contiguous()
    lib = _lib.load()
    outs, iters, resid = [], 0, 0.0
    R = _real(L.dtype)
    for c0 in range(0, rhs.shape[1], 4):
        rc = rhs[:, c0:c0 + 4].contiguous()
        c = rc.shape[1]
        work = torch.empty(int(lib.wiski_cg_work_elems(m, r, c)), dtype=L.dtype, device=L.device)
        x = torch.empty_like(rc)
        h_it, h_res = c_int(0), R(0)
        _call("wiski_cg_solve", L.dtype, _ptr(L), _ptr(KL), m, r, _ptr(rc), c, R(tol), int(max_iter), int(check_every),
              _ptr(x), ctypes.byref(h_it), ctypes.byref(h_res), _ptr(work), _stream())
        outs.append(x)
        iters, resid = max(iters, h_it.value), max(resid, h_res.value)
    return (outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)), iters, resid
