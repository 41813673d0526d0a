"""TEST-ONLY: replaces the CUDA entry points of ``online_gp_b200.ops`` with CPU restatements built on the oracle,
so that the *host logic* (LazyTensor algebra, cache protocol, model / MLL / wrapper orchestration, autograd
wiring) can be exercised without a GPU in the ``-m "not gpu"`` suite.  The product never uses this; the real
kernels are checked against the same oracle in the ``-m gpu`` tests."""
import contextlib

import torch

from oracle import gridkernel as ogk
from oracle import interp as oi


def _grid_from_spec(spec, dtype):
    # rebuild per-dim grids (float32 values) from the spec: lo + k * delta reproduces linspace only approximately,
    # so the mock keeps the original buffers attached by GridSpec.__init__ (see install()).
    return spec._grid_cpu


@contextlib.contextmanager
def install():
    from online_gp_b200 import ops

    saved = {k: getattr(ops, k) for k in ("_require_cuda", "_interp_fwd", "_gather", "_scatter_add", "_kron_mm",
                                          "_kron_bwd_cols", "_rmul", "_gram", "panel_lowrank_update_", "panel_lowrank_update2_", "panel_lowrank_update1_", "panel_outer_add_",
                                          "q_matvec",
                                          "cg_solve", "kron_axis_apply", "kron_axis_contract", "overlap_capable",
                                          "side_section", "join_side", "is_background", "background")}
    orig_init = ops.GridSpec.__init__
    orig_bwd = ops._InterpFn.backward

    def spec_init(self, grid):
        orig_init(self, grid)
        self._grid_cpu = [g.detach().cpu().clone() for g in grid]

    def interp_fwd(x, spec, check_bounds):
        if x.shape[0] == 0:
            return torch.zeros(0, spec.s, dtype=torch.long), torch.zeros(0, spec.s, dtype=x.dtype)
        idx, val = oi.interpolate(spec._grid_cpu, x.detach())
        return idx, val.to(x.dtype)

    def interp_bwd(ctx, _gidx, gval):
        (x,) = ctx.saved_tensors
        with torch.enable_grad():
            xx = x.detach().clone().requires_grad_(True)
            _, val = oi.interpolate(ctx.spec._grid_cpu, xx)
            (gx,) = torch.autograd.grad((val.to(gval.dtype) * gval).sum(), xx)
        return gx, None, None

    def gather(idx, val, src):
        return oi.left_interp(idx, val, src)

    def scatter_add(idx, val, src, dst):
        c = src.shape[-1]
        dst.index_add_(0, idx.reshape(-1), (val.unsqueeze(-1) * src.unsqueeze(1)).reshape(-1, c))
        return dst

    def kron_mm(cols, sizes, X):
        return ogk.kron_toeplitz_matmul([cols[i, :g] for i, g in enumerate(sizes)], X)

    def kron_bwd_cols(cols, sizes, Z, X):
        with torch.enable_grad():
            cc = cols.detach().clone().requires_grad_(True)
            out = ogk.kron_toeplitz_matmul([cc[i, :g] for i, g in enumerate(sizes)], X.detach())
            (g,) = torch.autograd.grad((out * Z.detach()).sum(), cc)
        return g

    def lowrank(P, U, Vt):
        P.add_((P @ U) @ Vt)
        return P

    def q_matvec(L, KL, v):
        return v + L.t() @ (KL @ v)

    def cg_solve(L, KL, rhs, tol=1e-2, max_iter=1000, check_every=4):
        Q = torch.eye(L.shape[1], dtype=L.dtype) + L.t() @ KL
        return torch.linalg.solve(Q, rhs), 1, 0.0

    def axis_apply(X, col, g, outer, inner):
        T = ogk.toeplitz_dense(col[:g])
        return torch.einsum("ab,obw->oaw", T, X.reshape(outer, g, inner)).reshape(X.shape).contiguous()

    def axis_contract(Z, P, g, outer, inner, acc64):
        G = torch.einsum("oaw,obw->ab", Z.reshape(outer, g, inner).double(), P.reshape(outer, g, inner).double())
        ar = torch.arange(g)
        off = (ar.unsqueeze(0) - ar.unsqueeze(1)).abs()
        acc64[:g] += torch.zeros(g, dtype=torch.float64).index_add_(0, off.reshape(-1), G.reshape(-1))
        return acc64

    ops.kron_axis_apply = axis_apply
    ops.kron_axis_contract = axis_contract
    ops.GridSpec.__init__ = spec_init
    ops._InterpFn.backward = staticmethod(interp_bwd)
    ops._require_cuda = lambda *a: None
    ops._interp_fwd = interp_fwd
    ops._gather = gather
    ops._scatter_add = scatter_add
    ops._kron_mm = kron_mm
    ops._kron_bwd_cols = kron_bwd_cols
    ops._rmul = lambda P, M, terms=3: P @ M
    ops._gram = lambda A, B, symmetric=False: A.t() @ B
    ops.panel_lowrank_update_ = lowrank
    def lowrank2(P0, P1, U, Vt0, Vt1, return_t=False):
        T = P0 @ U
        out = (lowrank(P0, U, Vt0), lowrank(P1, U, Vt1))
        return out + (T,) if return_t else out
    ops.panel_lowrank_update2_ = lowrank2
    def lowrank1(P, U, Vt, return_t=False):
        T = P @ U
        lowrank(P, U, Vt)
        return (P, T) if return_t else P
    ops.panel_lowrank_update1_ = lowrank1
    ops.panel_outer_add_ = lambda P, T, W: P.add_(T @ W)
    ops.q_matvec = q_matvec
    ops.cg_solve = cg_solve

    # settings.overlap_root_update: the side-stream schedule runs inline on the CPU (same order of operations, no streams)
    class _Inline:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    ops.overlap_capable = lambda t: True
    ops.side_section = _Inline
    ops.background = _Inline
    ops.join_side = lambda device: None
    ops.is_background = lambda: False
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
        ops.GridSpec.__init__ = orig_init
        ops._InterpFn.backward = orig_bwd
